/* lrm_c.h — C ABI of the B200-native leg-movability library (liblrm_b200.so).
 *
 * This is the drop-in boundary for the reference's host<->CUDA plugin surface.  The reference has
 * no extern "C" layer: plain-C++ callers (bench.cpp, several_leg.cpp) reach CUDA through the C++
 * symbols declared in cross_compiled.cuh / one_leg.cu.h / several_leg.cu.h /
 * several_leg_octree.cu.h.  Each entry point below names the reference interface it replaces;
 * include/lrm_compat.hpp re-exports the reference's own C++ signatures on top of these.
 *
 * Conventions
 *   - units: millimetres and radians (static_variables.cpp:45-49), all arithmetic FP32.
 *   - points / vectors are N x 3 float AoS (the layout of the reference's Array<float3>).
 *   - flags are one byte per element, 0 or 1 (the layout of Array<bool>).
 *   - quat is the body orientation in the reference's storage (settings.h:51 quatTest,
 *     unified_math_cuda.cu.h:13-27): {1,0,0,0} is the identity.  NULL means identity.
 *   - on_device != 0: all data pointers are device pointers on the current device.  The one-leg
 *     entry points (lrm_reach, lrm_dist, lrm_reach_dist(_soa), lrm_forward_kine, lrm_recurs,
 *     lrm_make_lattice) are asynchronous on `stream` unless kernel_ms is requested (then they
 *     synchronise the stream) — also on the first call of a new (leg, orientation), whose certified
 *     tables are built on `stream`, and when a cached plan is evicted (the rebuild is ordered behind
 *     the victim's recorded uses by events).  lrm_positionability, lrm_oct* build per-call scratch
 *     (the cell grid of the map) and return when their work on `stream` has finished.
 *     on_device == 0: pointers are host pointers; the library stages through device memory
 *     (alloc, H2D, kernel, D2H, free — the contract of apply_kernel, cross_compiled.cu:34-79).
 *   - out_xyz may be the very buffer xyz (an in-place call): every sweep reads a point before it
 *     writes that point's vector, and points that are finished late (the rings of the large sweeps)
 *     are re-read from an input their tile's store has left unchanged.  Partially overlapping
 *     buffers are not supported.
 *   - kernel_ms (may be NULL) receives the cudaEvent time of the kernel(s) only — the number the
 *     reference's apply_kernel returns (cross_compiled.cu:58-65).
 *   - return value: 0 on success, negative lrm_status otherwise; lrm_last_error() gives the text.
 *     No entry point ever computes on the CPU: without a usable CUDA device every compute call
 *     fails with LRM_ERR_CUDA.
 */
#ifndef LRM_C_H
#define LRM_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRM_ABI_VERSION 1

#if defined(__GNUC__)
#define LRM_API __attribute__((visibility("default")))
#else
#define LRM_API
#endif

typedef enum {
    LRM_OK = 0,
    LRM_ERR_INVALID = -1, /* bad argument (NULL pointer, misaligned buffer, bad count) */
    LRM_ERR_CUDA = -2,    /* CUDA runtime error, text in lrm_last_error() */
    LRM_ERR_UNSUPPORTED = -3
} lrm_status;

/* Field-for-field the reference's LegDimensions (HeaderCPP.h:19-52): 14 floats, 56 bytes. */
typedef struct lrm_leg {
    float body_angle;
    float body;
    float coxa_pitch;
    float coxa_length;
    float tibia_length;
    float femur_length;
    float tibia_absolute_pos;
    float tibia_absolute_neg;
    float max_angle_coxa;
    float min_angle_coxa;
    float max_angle_tibia;
    float min_angle_tibia;
    float max_angle_femur;
    float min_angle_femur;
} lrm_leg_t;

/* ---- library / device ------------------------------------------------------------------- */
LRM_API int lrm_abi_version(void);
LRM_API const char* lrm_last_error(void);
/* Tuning knob: one-leg sweeps of at least n points go through the certified tables (plane atlas,
 * yaw sectors; same results).  Default 4 Mi points — building the 16 MiB atlas of
 * a new (leg, orientation) costs about 0.3 ms.  Returns the previous value. */
LRM_API size_t lrm_set_fast_path_min_points(size_t n);
/* Tuning / measurement knobs, process-wide, none of which can change a result (every sweep returns
 * the same bits).  *previous (may be NULL) receives the old value.
 *   "fast_path_min_points"  see lrm_set_fast_path_min_points
 *   "sweep"                 which sweep large distance calls take: 0 two-tier (certified tables +
 *                           full evaluation), 1 tiered (choice volume), 2 (default) chosen per launch
 *                           on the device by a coherence probe of the input
 *   "tier_chunk_shift"      log2 of the consecutive tiles a CTA of the tiered sweep takes (default 3)
 *   "volume_cell_mm", "volume_dim"   cube size and cubes per side of choice volumes built from now
 *                           on (default 3 mm x 512: 537 MB per cached plan)
 *   "volume_bricks"         1 (default): cubes the grid cannot settle carry a brick of 4^3 fine cubes
 *                           (up to 1.4 GB more per cached plan at the default shape); 0: coarse grid only
 *   "staging_chunk_points"  points per chunk of the host-pointer pipeline (default 2 Mi)
 *   "skeleton"              measurement builds only (LRM_ERR_UNSUPPORTED otherwise) */
LRM_API int lrm_set_option(const char* name, double value, double* previous);
/* Counters: "table_builds" (plane-atlas builds since the library was loaded: a cached plan builds
 * nothing), "volume_builds" (background builds of a choice volume that a later sweep found finished),
 * "volume_cell_mm", "volume_dim", "volume_bricks" / "volume_brick_capacity" (bricks the last finished
 * volume build asked for / the size of its pool). */
LRM_API int lrm_get_stat(const char* name, double* value);
LRM_API int lrm_device_count(void);
LRM_API int lrm_set_device(int device);

/* Default legs of the reference: robot 0 = get_moonbot_leg, 1 = get_M2_leg
 * (static_variables.cpp:44-93).  Pure host arithmetic. */
LRM_API int lrm_default_leg(int robot, float azimuth, lrm_leg_t* out);

/* ---- one-leg hot path -------------------------------------------------------------------- */
/* Replaces apply_kernel(points, dim, reachability_global_kernel, out)
 * (cross_compiled.cuh:4-7, one_leg_global.cu:149-156; one_leg.cu:343-357 when quat is identity). */
LRM_API int lrm_reach(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, uint8_t* flags,
              int on_device, void* stream, float* kernel_ms);

/* Replaces apply_kernel(points, dim, distance_global_kernel, out)
 * (one_leg_global.cu:157-166; one_leg.cu:359-375).  out_xyz[i] = d(p_i) with p_i - d(p_i) on the
 * reachability edge.  flags (may be NULL) receives distance_global's bool. */
LRM_API int lrm_dist(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, float* out_xyz,
             uint8_t* flags, int on_device, void* stream, float* kernel_ms);

/* Both results in one pass over the points (25 B/point of HBM traffic instead of 13 + 24):
 * reach_flags[i] == reachability_global(p_i), out_xyz[i] == distance_global's vector. */
LRM_API int lrm_reach_dist(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat,
                   uint8_t* reach_flags, float* out_xyz, int on_device, void* stream,
                   float* kernel_ms);

/* SoA twins: x, y, z planes in / dx, dy, dz planes out (the layout of the reference's on-disk
 * protocol dist_input_t{x,y,z}.bin -> out_dist_x{x,y,z}.bin, several_leg.cpp:126-221).
 * dx/dy/dz may all be NULL to skip the vectors.  Device pointers only, every plane 16-byte aligned
 * (the bulk-copy engine's requirement; misaligned planes are refused with LRM_ERR_CUDA). */
LRM_API int lrm_reach_dist_soa(const float* x, const float* y, const float* z, size_t n,
                       const lrm_leg_t* leg, const float* quat, uint8_t* reach_flags, float* dx,
                       float* dy, float* dz, void* stream, float* kernel_ms);

/* Forward kinematics (coxa, femur, tibia) -> xyz, replaces forward_kine_kernel
 * (one_leg.cu:377-414; like the reference it ignores coxa_pitch). */
LRM_API int lrm_forward_kine(const float* angles, size_t n, const lrm_leg_t* leg, float* out_xyz,
                     int on_device, void* stream, float* kernel_ms);

/* ---- benchmark support ------------------------------------------------------------------- */
/* Regular lattice written straight into device memory, x-major / z-fastest like
 * generate3DGrid (bench.cpp:30-50): point i -> (ix, iy, iz) = (i / (ny*nz), (i / nz) % ny, i % nz),
 * coordinate = lo + (float)index * step, evaluated with separately rounded multiply and add so
 * that a host loop doing the same two float operations is bit-identical.
 * Writes points [first, first + count) to out_xyz (device, AoS). */
LRM_API int lrm_make_lattice(float* out_xyz, const float lo[3], const float step[3], const uint32_t dims[3],
                     size_t first, size_t count, void* stream);

/* ---- multi-leg body positionability ------------------------------------------------------ */
typedef struct lrm_posit_opts {
    int pre_cull;        /* apply the constructor culls of multi_rot_estimator
                            (several_leg.cu:371-374): 60 mm colliding sphere, 400 mm far body,
                            400 mm far target.  robot_full_struct always does. */
    int first_hit_only;  /* reserved, must be 0 */
} lrm_posit_opts_t;

/* Orientation set of robot_full_struct (several_leg.cu:811-857): writes 45 quaternions. */
LRM_API int lrm_full_struct_orientations(float* out_quat4, int capacity);
/* RPYtoQuat (octree_util.cu.h:164-172): roll, then pitch, then yaw, in the storage qtRotate
 * expects — the way to build custom orientation sets (e.g. a yaw grid) for lrm_positionability. */
LRM_API int lrm_rpy_to_quat(float roll, float pitch, float yaw, float out_quat4[4]);

/* Replaces robot_full_struct's pipeline (several_leg.cu:326-877) with a per-pose result instead
 * of a compacted list: standable[b] = 1 + index of the first orientation in `quats` for which the
 * body position bodies[b] passes the cull cylinders and EVERY leg reaches at least one map point
 * (reachable_rotate_leg, several_leg.cu:48-67); 0 if none.  The reference hard-codes 4 legs
 * (several_leg.cu:681-697); nlegs is free here (hexapod = 6).
 * bodies: nb x 3, map: nt x 3 (device pointers when on_device).  Unlike the one-leg sweeps this
 * call returns only when the search has finished on `stream` (its per-call scratch — the cell grid
 * of the map, the plans — is released on return). */
LRM_API int lrm_positionability(const float* bodies, size_t nb, const float* map, size_t nt,
                        const lrm_leg_t* legs, int nlegs, const float* quats, int nq,
                        const lrm_posit_opts_t* opts, uint8_t* standable, int on_device,
                        void* stream, float* kernel_ms);

/* The same search with its work counted (an instrumented instantiation of the kernel: not for
 * timing).  counts[0] = leg predicates executed (reachable_rotate_leg, several_leg.cu:48-67, on one
 * map point), counts[1] = cull-cylinder predicates executed (several_leg.cu:504-559), counts[2] =
 * the ALGORITHMIC leg-predicate count of these poses: for every orientation, every map
 * point inside the reach cylinder, once per leg — what reach_mem_kernel (several_leg.cu:92-129)
 * evaluates without pruning or early exit.  standable is filled as by lrm_positionability. */
LRM_API int lrm_positionability_counts(const float* bodies, size_t nb, const float* map, size_t nt,
                               const lrm_leg_t* legs, int nlegs, const float* quats, int nq,
                               const lrm_posit_opts_t* opts, uint8_t* standable, double counts[3],
                               int on_device, void* stream);

/* Replaces apply_recurs<float3,LegDimensions,float3> (cross_compiled.cuh:9-10, cross_compiled.cu:82-139;
 * recursive_kernel one_leg_global.cu:168-251, fillOutKernel octree_util.cu:9-26): adaptive octree of
 * the single-leg distance field (root box +-5000 mm, MINBOXSIZE 100 mm), refined to `max_depth`
 * (the reference's compile-time MAX_DEPTH), painted on the query points: out_xyz[i] =
 * (depth of the leaf box containing p_i, 0, 0).  Points outside the root box are left untouched
 * (the reference leaves uninitialised device memory there). */
LRM_API int lrm_recurs(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, int max_depth,
               float* out_xyz, int on_device, void* stream, float* kernel_ms);

/* ---- body-space octree ------------------------------------------------------------------- */
/* Replaces apply_oct (several_leg_octree.cu.h:4, several_leg_octree.cu:391-488): adaptive octree
 * over BODY positions (root box +-5000 mm, settings.h:26; axes stop splitting below 100 mm,
 * settings.h:17), refined `max_depth` times (the reference's compile-time MAX_DEPTH, settings.h:15,
 * ships as 1).  A child node is valid when, for some foothold and orientation sample
 * (octree_util.cu.h:184-198), all four legs mounted at k*pi/4 (settings.h:41-42) reach that same
 * foothold; it stays "on edge" (and is refined next pass) when a leg's distance vector falls inside
 * the child box.  Semantics are those of a sequential evaluation (flags OR-ed over all work items
 * of a pass): the reference's own GPU result is racy (unsynchronised shared flags).
 * Writes the centres of the valid leaf / raw nodes in the reference's traversal order
 * (octree_util.cu:123-147) to out_xyz (host, capacity `cap` points) and their number to *count
 * (which may exceed cap: call again with a larger buffer).  footholds: host or device pointer. */
LRM_API int lrm_oct(const float* footholds, size_t nt, const lrm_leg_t* leg, int max_depth, float* out_xyz,
            size_t cap, size_t* count, int on_device, void* stream, float* kernel_ms);

/* lrm_oct for ONE shard of the tree: the subtrees under the root's eight children never interact
 * (branchKernel descends child by child, several_leg_octree.cu:296-313), so shard `shard` of
 * `nshards` refines only the top-level children c with c % nshards == shard — one process per GPU,
 * footholds replicated, no collective.  child_counts[c] (8 entries, may be NULL) receives the number
 * of centres written for top-level child c; out_xyz holds them grouped by ascending c, so the
 * full result in the reference's traversal order is, for c = 0..7, shard (c % nshards)'s group c. */
LRM_API int lrm_oct_sharded(const float* footholds, size_t nt, const lrm_leg_t* leg, int max_depth, int shard,
                    int nshards, float* out_xyz, size_t cap, size_t* count, size_t child_counts[8],
                    int on_device, void* stream, float* kernel_ms);

/* Replaces ONE launch of validity_child (several_leg_octree.cu:19-151) together with the child
 * initialisation branchKernel does before it (:315-352): the eight children of the body box
 * parent_box6 = {centre xyz, half extents xyz} (parent_validity = the parent's validity flag, which
 * the predicate inherits, :71) are created with CreateChildBox's rules and evaluated against all
 * footholds.  out_flags: 8 x {validity, leaf, raw, onEdge} bytes; out_boxes: 8 x 6 floats (both
 * host).  Sequential semantics, like lrm_oct. */
LRM_API int lrm_oct_children(const float* footholds, size_t nt, const lrm_leg_t* leg, const float parent_box6[6],
                     int parent_validity, uint8_t* out_flags, float* out_boxes, int on_device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LRM_C_H */
