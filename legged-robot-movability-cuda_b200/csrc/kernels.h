// kernels.h — launchers shared between the kernel translation units and the C ABI (lrm_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "leg_plan.h"

namespace lrm {

enum : int { kModeReach = 1, kModeDist = 2, kModeBoth = 3 };

// One-leg sweep over N x 3 AoS points already on the device.  mode: kModeReach -> flag only,
// kModeDist -> vector (+ distance_global's bool if flag != nullptr), kModeBoth -> vector +
// reachability flag.
// n_call: size of the whole call this launch is a chunk of (0 = n); the certified tables of the
// distance fast path are built / used when the call, not the chunk, is large enough to pay for them.
cudaError_t launch_one_leg_aos(int mode, const LegPlan& plan, const float* xyz, float* out_vec,
                               uint8_t* flag, size_t n, cudaStream_t stream, size_t n_call = 0);
// Sweeps (calls) of at least this many points use the certified tables (default 4 Mi: building the
// atlas costs ~0.3 ms per new plan); returns the previous value.
size_t set_fast_path_min_points(size_t n);
// lrm_set_option: 0 = two-tier sweep, 1 = tiered sweep, 2 = chosen per launch by the coherence probe;
// log2 of the tiles per chunk of the tiered sweep; the staging skeleton (measurement builds only, else -1).
// Each returns the previous value.
int set_sweep_mode(int mode);
int set_tier_chunk_shift(int shift);
int set_tier_kernel(int which);  // 0 = CTA-tiled tiered sweep, 1 = warp-autonomous tiered sweep
int set_skeleton(int on);
// SoA planes; dx == nullptr selects reach-only.
cudaError_t launch_one_leg_soa(const LegPlan& plan, const float* x, const float* y, const float* z,
                               float* dx, float* dy, float* dz, uint8_t* flag, size_t n,
                               cudaStream_t stream);
// apply_recurs: (octree leaf depth, 0, 0) per point; points outside the +-5000 mm root box are
// not written.
cudaError_t launch_recurs(const LegPlan& plan, const float* xyz, float* out, size_t n, int max_depth,
                          cudaStream_t stream);
cudaError_t launch_forward_kine(const float* angles, const lrm_leg_t& leg, float* out, size_t n,
                                cudaStream_t stream);
cudaError_t launch_lattice(float* out, const float lo[3], const float step[3],
                           const uint32_t dims[3], size_t first, size_t count,
                           cudaStream_t stream);

// Certified tables of a plan's distance fast path (plane_atlas.cu): the plane atlas (device) and
// the yaw-sector table (host, passed as a kernel parameter); built on first use, cached per device
// (8 plans each).  acquire_tables PINS the entry: it cannot be evicted or rebuilt until
// release_tables, which must follow the caller's last launch that reads the views (it records that
// use on `stream`, so that a later rebuild is ordered after it on the device).  No host wait.
struct AtlasView;
struct TableLease {
    void* entry = nullptr;
};
cudaError_t acquire_tables(const LegPlan& plan, cudaStream_t stream, AtlasView* view, FastTables* tables,
                           TableLease* lease);
void release_tables(TableLease* lease, cudaStream_t stream);
// true if acquire_tables(plan) would build nothing on the current device
bool tables_cached(const LegPlan& plan);

// Choice volume of a leased plan: 3-D texture of 32-bit texels (winning coxa solution + plane label
// per cube, or a pointer to a brick of 4^3 fine cubes, plane_atlas.cu).  Built in the background on first request: until it is there the call
// returns cudaErrorNotReady (wait = false) or blocks (wait = true).
struct VolumeView;
cudaError_t get_choice_volume(const TableLease& lease, cudaStream_t stream, VolumeView* view, bool wait);
// cube size (mm) and cubes per side (multiple of 4) of volumes built from now on; -1 if out of range
int set_choice_volume_shape(float cell_mm, int dim);
void get_choice_volume_shape(float* cell_mm, int* dim);
int set_volume_bricks(int on);                               // bricks under uncertified cubes; returns the previous value
void get_brick_stats(unsigned* used, unsigned* capacity);    // of the last finished volume build
unsigned long long table_builds();  // atlas builds since load (diagnostics)
unsigned long long volume_builds_done();  // choice-volume builds a sweep has seen finished since load

// Multi-leg positionability (positionability.cu).  All pointers are device pointers.
struct PositParams {
    const float* bodies;   // nb x 3
    size_t nb;
    const float* map;      // nt x 3
    size_t nt;
    const lrm_leg_t* legs; // host
    int nlegs;
    const float* quats;    // host, nq x 4
    int nq;
    int pre_cull;
    uint8_t* standable;    // nb
    double* stats = nullptr;  // host, 3 entries, or nullptr: see lrm_positionability_counts
};
cudaError_t run_positionability(const PositParams& p, cudaStream_t stream, float* kernel_ms);

// Body-space octree (octree.cu).  d_footholds: device, nt x 3.  centres: xyz triples of the valid
// leaf / raw nodes in the reference's traversal order.
// shard / nshards: only the root's children c with c % nshards == shard are refined (the subtrees
// are independent); child_counts (8, may be nullptr): centres found under each top-level child.
cudaError_t run_octree(const float* d_footholds, size_t nt, const lrm_leg_t& leg, int max_depth,
                       std::vector<float>* centres, cudaStream_t stream, float* kernel_ms, int shard = 0,
                       int nshards = 1, size_t* child_counts = nullptr);
// One pass of validity_child over the 8 children of one parent box: flags32 = 8 x {validity, leaf,
// raw, onEdge}, boxes48 = 8 x {centre, half extents} (host).
cudaError_t run_octree_children(const float* d_footholds, size_t nt, const lrm_leg_t& leg, const float* parent_box6,
                                int parent_validity, uint8_t* flags32, float* boxes48, cudaStream_t stream);

}  // namespace lrm
