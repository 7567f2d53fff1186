// one_leg_kernels.cu — streaming kernels of the one-leg sweep (reach / distance / both).
//
// Replaces reachability_global_kernel, distance_global_kernel, reachability_circles_kernel,
// distance_circles_kernel (one_leg_global.cu:149-166, one_leg.cu:343-375): one thread per point,
// <<<ceil(N/256),256>>>, AoS float3 LDG/STG with 12-byte stride and a Circle[14] stack frame.
//
// Here: a persistent grid (a few CTAs per SM) walks the point array in tiles.  Tiles are staged
// global -> shared by the bulk-copy engine (cp.async.bulk, mbarrier completion, kStages deep),
// threads read their points from shared memory with conflict-free strides, keep every
// intermediate in registers, write results back to shared memory (in place) and one thread per
// CTA sends the tile to global memory with a bulk store.  Algorithmic HBM traffic: 12 B in + 1 B (flag)
// and/or 12 B (vector) out per point — each byte crosses HBM exactly once.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "bulk_copy.cuh"
#include "kernels.h"
#include "leg_math.cuh"

namespace lrm {

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 1024;  // points per tile (multiple of 16 keeps every bulk copy 16-B sized)
constexpr int kPrefetch = 2;  // tiles in flight ahead of the compute
#ifndef LRM_SKIP_B
#define LRM_SKIP_B 0
#endif
constexpr bool kSkipB = LRM_SKIP_B != 0;  // skip the flipped solution by its lower bound: a branch that
// keeps the two points of a trip from interleaving (measured: 104 vs 111 Gpoints/s) -> off
constexpr size_t kAtlasMinPoints = size_t(1) << 22;
std::atomic<size_t> g_fast_min_points{kAtlasMinPoints};
// ... but once a plan's tables are cached they cost nothing: sweeps from this size on use them
constexpr size_t kCachedMinPoints = size_t(1) << 20;  // (measured on the reference's size sweep: the probe +
// two gated launches cost more than they save at 256 Ki points)
// below this, a launch is latency: the plain kernel (no staging pipeline to fill and drain)
constexpr size_t kPlainMaxPoints = size_t(1) << 14;
// deferred points of the fast path: entries of at most two tiles plus a partial flush block
constexpr int kQueueCap = 2 * kTile + kThreads + 256;
static_assert(kTile % kThreads == 0 && kTile % 16 == 0 && kTile <= 1024, "tile shape");

// shared-memory state of the fast path: certified tables + the ring of parked points
struct FastSmem {
    alignas(16) WinnerTable winners;
    alignas(16) YawPair ypair[kYawPairs];
    alignas(16) unsigned char ycode[kYawBins + 16];
    uint32_t queue[kQueueCap];  // ring: iteration << 10 | index in tile
    unsigned qcnt[3];           // entries appended in iteration it % 3
};
struct NoFastSmem {};

template <int MODE, bool FAST>
struct alignas(128) StreamSmem {
    // Distance modes work IN PLACE: a thread overwrites its point with the point's vector (same 12
    // bytes), and the tile goes back to global memory from the buffer it arrived in.  Three
    // buffers rotate (being loaded / being computed / being stored); reach-only needs two.  (Two
    // buffers for the distance modes as well, as in the tiered sweep: 115.2 vs 114.0 Gpoints/s — not kept.)
    static constexpr int kStages = (MODE & kModeDist) ? 3 : 2;
    float in[kStages][3 * kTile];
    uint8_t flag[2][kTile];
    alignas(16) SectorTable table;
    alignas(16) typename std::conditional<(FAST && (MODE & kModeDist) != 0), FastSmem, NoFastSmem>::type fast;
    alignas(8) uint64_t full[kStages];
};

// One point through the full evaluation.
template <int MODE, bool SOA, bool GENERIC>
__device__ __forceinline__ void compute_point(const LegPlan& L, const SectorTable& tab,
                                              const float* in, float* vec, uint8_t* flag, int i,
                                              int stride_pts) {
    float x, y, z;
    if (SOA) {
        x = in[i], y = in[stride_pts + i], z = in[2 * stride_pts + i];
    } else {
        x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];  // word stride 3: conflict-free
    }
    const CoxaPoint p = to_coxa_frame(L, x, y, z);
    if (MODE == kModeReach) {
        flag[i] = reach_coxa_frame(L, tab, p) ? 1 : 0;
    } else {
        const DistResult r = dist_coxa_frame<GENERIC>(L, tab, p);
        if (SOA) {
            vec[i] = r.dx, vec[stride_pts + i] = r.dy, vec[2 * stride_pts + i] = r.dz;
        } else {
            vec[3 * i] = r.dx, vec[3 * i + 1] = r.dy, vec[3 * i + 2] = r.dz;
        }
        // MODE dist: distance_global's bool; MODE both: reachability_global's bool
        flag[i] = (MODE == kModeBoth ? r.reach : r.flag) ? 1 : 0;
    }
}

// One point through the certified tables; false (nothing written) when they cannot decide it.
template <int MODE, bool SOA, bool TEX>
__device__ __forceinline__ bool compute_point_fast(const LegPlan& L, const FastView& F,
                                                   const AtlasView& A, const WinnerTable& W,
                                                   const float* in, float* vec, uint8_t* flag, int i) {
    float x, y, z;
    if (SOA) {
        x = in[i], y = in[kTile + i], z = in[2 * kTile + i];
    } else {
        x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
    }
    DistResult r;
    const bool ok = dist_fast<TEX, kSkipB>(L, F, A, W, to_coxa_frame(L, x, y, z), &r);
    // an uncertified point leaves its input in place (values are garbage then; the redo rewrites
    // the slot in global memory after the tile's store)
    if (ok) {
        if (SOA) {
            vec[i] = r.dx, vec[kTile + i] = r.dy, vec[2 * kTile + i] = r.dz;
        } else {
            vec[3 * i] = r.dx, vec[3 * i + 1] = r.dy, vec[3 * i + 2] = r.dz;
        }
        flag[i] = (MODE == kModeBoth ? r.reach : r.flag) ? 1 : 0;
    }
    return ok;
}

// The full evaluation of one point straight from / to global memory.  Deliberately NOT inlined
// into the streaming loop: inlined, the compiler hoists this path's ~60 plan constants into
// uniform registers at the top of every tile whether or not anything is redone.
template <int MODE, bool SOA>
__device__ __noinline__ void redo_point(const LegPlan& L, const SectorTable& tab,
                                        const float* __restrict__ in_x, const float* __restrict__ in_y,
                                        const float* __restrict__ in_z, float* __restrict__ out_x,
                                        float* __restrict__ out_y, float* __restrict__ out_z,
                                        uint8_t* __restrict__ out_flag, size_t g) {
    float xyz[3], v[3];
    uint8_t f;
    if (SOA) {
        xyz[0] = in_x[g], xyz[1] = in_y[g], xyz[2] = in_z[g];
    } else {
        xyz[0] = in_x[3 * g], xyz[1] = in_x[3 * g + 1], xyz[2] = in_x[3 * g + 2];
    }
    compute_point<MODE, false, false>(L, tab, xyz, v, &f, 0, 0);
    if (SOA) {
        out_x[g] = v[0], out_y[g] = v[1], out_z[g] = v[2];
    } else {
        out_x[3 * g] = v[0], out_x[3 * g + 1] = v[1], out_x[3 * g + 2] = v[2];
    }
    if (out_flag) out_flag[g] = f;
}

// AoS: in_x = xyz (N x 3), out_x = vectors (N x 3).  SoA: separate planes.
//
// FAST (distance modes, standard legs, large sweeps): every point first goes through the
// certified tables (dist_fast).  The few points they cannot decide are NOT redone inside the
// tile — that would leave most warps idle at the tile barrier while two or three of them run the
// long evaluation — but parked in a per-CTA ring and redone later, 256 at a time by all threads
// (dense warps, equal work), straight from / to global memory.  A parked point is only redone
// once the bulk store of its tile has completed (its slot in the output then holds stale bytes
// that the late plain store replaces): thread 0 waits for all committed stores before every tile
// barrier, so entries appended two iterations ago are safe.
template <int MODE, bool SOA, bool GENERIC, bool FAST, bool TEX>
__global__ void __launch_bounds__(kThreads)
    one_leg_stream_kernel(const __grid_constant__ LegPlan L, const __grid_constant__ FastTables FT,
                          const AtlasView atlas, const float* __restrict__ in_x,
                          const float* __restrict__ in_y, const float* __restrict__ in_z,
                          float* __restrict__ out_x, float* __restrict__ out_y,
                          float* __restrict__ out_z, uint8_t* __restrict__ out_flag, size_t n,
                          const int* __restrict__ gate, int gate_want, const VolumeView vol) {
    if (gate != nullptr && gate_want >= 0 && *gate != gate_want) return;  // the coherence probe chose the other sweep
    // reach-only: gate_want < 0 means "read the reach bits of the choice volume if the probe found
    // the input coherent"; vol.tex == 0: no volume (yet)
    const bool reach_vol = vol.tex != 0 && (gate == nullptr || *gate == 1);
    // keep the pointer provably in the shared window (LDS/STS, not generic LD/ST): no integer
    // round-trip on the address; the bulk engine only needs 16-byte alignment
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& S = *reinterpret_cast<StreamSmem<MODE, FAST>*>(smem_raw);
    const int tid = threadIdx.x;

    const size_t n_bulk = n & ~size_t(15);  // points that move through the bulk engine
    const size_t n_tiles = (n_bulk + kTile - 1) / kTile;
    constexpr bool kVec = (MODE & kModeDist) != 0;
    constexpr bool kRedo = FAST && kVec;       // distance fast path: certified tables + deferred redo
    constexpr bool kReachAtlas = FAST && !kVec;  // reach-only: valid bit of the atlas cell

    fill_sector_table(L, &S.table, tid, kThreads);
    if constexpr (kRedo) {
        fill_winner_table(L, &S.fast.winners, tid, kThreads);
        for (int i = tid; i < kYawPairs * (int)(sizeof(YawPair) / 4); i += kThreads)
            reinterpret_cast<float*>(S.fast.ypair)[i] = reinterpret_cast<const float*>(FT.pair)[i];
        for (int i = tid; i < (kYawBins + 16) / 4; i += kThreads)
            reinterpret_cast<uint32_t*>(S.fast.ycode)[i] = reinterpret_cast<const uint32_t*>(FT.code)[i];
        if (tid == 0) S.fast.qcnt[0] = S.fast.qcnt[1] = S.fast.qcnt[2] = 0;
    }
    constexpr int kStages = StreamSmem<MODE, FAST>::kStages;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) bulk::mbar_init(&S.full[s], 1);
        bulk::fence_barrier_init();
    }
    __syncthreads();

    auto tile_count = [&](size_t tile) -> uint32_t {
        const size_t first = tile * kTile;
        return (uint32_t)((n_bulk - first < (size_t)kTile) ? (n_bulk - first) : kTile);
    };
    auto issue_load = [&](size_t tile, int stage) {
        const uint32_t cnt = tile_count(tile);
        const size_t first = tile * kTile;
        if (SOA) {
            bulk::mbar_expect_tx(&S.full[stage], 3 * cnt * 4);
            bulk::load(&S.in[stage][0], in_x + first, cnt * 4, &S.full[stage]);
            bulk::load(&S.in[stage][kTile], in_y + first, cnt * 4, &S.full[stage]);
            bulk::load(&S.in[stage][2 * kTile], in_z + first, cnt * 4, &S.full[stage]);
        } else {
            bulk::mbar_expect_tx(&S.full[stage], cnt * 12);
            bulk::load(&S.in[stage][0], in_x + 3 * first, cnt * 12, &S.full[stage]);
        }
    };
    // a parked point, redone from / to global memory
    auto redo = [&](uint32_t entry, uint32_t it_now) {
        const uint32_t age = (it_now - (entry >> 10)) & 0x3fffffu;
        const size_t g = ((size_t)blockIdx.x + (size_t)(it_now - age) * gridDim.x) * kTile + (entry & 1023u);
        redo_point<MODE, SOA>(L, S.table, in_x, in_y, in_z, out_x, out_y, out_z, out_flag, g);
    };

    if (tid == 0) {
        for (int s = 0; s < kPrefetch; s++) {
            const size_t tile = (size_t)blockIdx.x + (size_t)s * gridDim.x;
            if (tile < n_tiles) issue_load(tile, s);
        }
    }

    FastView fview{nullptr, nullptr, nullptr, 0};
    if constexpr (kRedo) fview = FastView{S.fast.ypair, S.fast.ycode, nullptr, 0};
    uint32_t it = 0;
    // ring bookkeeping, identical in every thread: entries appended before this iteration (base),
    // before the previous one (elig: their tiles' stores have completed), and redone so far (head)
    uint32_t q_base = 0, q_elig = 0, q_head = 0;
    int rot = 0;  // it % 3
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = it % kStages;
        const int ob = it & 1;
        const uint32_t cnt = tile_count(tile);
        bulk::mbar_wait(&S.full[stage], (it / kStages) & 1);

        float* in = S.in[stage];
        float* vec = in;  // in place (unused in reach-only mode)
        uint8_t* flag = S.flag[ob];
        if constexpr (kReachAtlas) {
            if (reach_vol && cnt == (uint32_t)kTile) {
                // The cube of a point may be reachable, or unreachable, as a whole (reach bits of
                // the choice volume): such a point needs nothing but its cube byte.  All of a
                // thread's volume fetches are in flight together; the few points in undecided
                // cubes take the yaw tests + atlas cell below.
                constexpr int kPer = kTile / kThreads;
                CoxaPoint p[kPer];
                unsigned c[kPer];
#pragma unroll
                for (int k = 0; k < kPer; k++) {
                    const int i = tid + k * kThreads;
                    float x, y, z;
                    if (SOA) {
                        x = in[i], y = in[kTile + i], z = in[2 * kTile + i];
                    } else {
                        x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
                    }
                    p[k] = to_coxa_frame(L, x, y, z);
                    // the reach bits sit in bits 5-6 of either kind of texel (classic or brick pointer)
                    c[k] = tex3D<unsigned>(vol.tex, fmaf(p[k].x, vol.inv_cell, vol.o),
                                           fmaf(p[k].y, vol.inv_cell, vol.oy), fmaf(p[k].z, vol.inv_cell, vol.o));
                }
#pragma unroll
                for (int k = 0; k < kPer; k++) {
                    const bool r = (c[k] & kVolReachKnown) ? (c[k] & kVolReachValue) != 0u
                                                           : reach_coxa_frame_atlas<TEX>(L, S.table, atlas, p[k]);
                    flag[tid + k * kThreads] = r ? 1 : 0;
                }
            } else
#pragma unroll 2
            for (int i = tid; i < (int)cnt; i += kThreads) {
                float x, y, z;
                if (SOA) {
                    x = in[i], y = in[kTile + i], z = in[2 * kTile + i];
                } else {
                    x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
                }
                flag[i] = reach_coxa_frame_atlas<TEX>(L, S.table, atlas, to_coxa_frame(L, x, y, z)) ? 1 : 0;
            }
        } else if constexpr (kRedo) {
            const int rot_prev = rot == 0 ? 2 : rot - 1;
            q_elig = q_base;
            q_base += S.fast.qcnt[rot_prev];  // final since the previous tile barrier
#pragma unroll 1
            while (q_elig - q_head >= (uint32_t)kThreads) {
                redo(S.fast.queue[(q_head + tid) % kQueueCap], it);
                q_head += kThreads;
            }
            // two points per trip, both computed before either is parked: the fast path is
            // straight-line code, so the two evaluations (and their texture fetches) interleave
            auto park = [&](int i) {
                const uint32_t pos = q_base + atomicAdd(&S.fast.qcnt[rot], 1u);
                S.fast.queue[pos % kQueueCap] = (it << 10) | (uint32_t)i;
            };
            auto load_pt = [&](int i, float& x, float& y, float& z) {
                if (SOA) {
                    x = in[i], y = in[kTile + i], z = in[2 * kTile + i];
                } else {
                    x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
                }
            };
            // an uncertified point leaves its slot alone (the redo rewrites it in global memory
            // after the tile's store)
            auto store_pt = [&](int i, const DistResult& r) {
                if (SOA) {
                    vec[i] = r.dx, vec[kTile + i] = r.dy, vec[2 * kTile + i] = r.dz;
                } else {
                    vec[3 * i] = r.dx, vec[3 * i + 1] = r.dy, vec[3 * i + 2] = r.dz;
                }
                flag[i] = (MODE == kModeBoth ? r.reach : r.flag) ? 1 : 0;
            };
#pragma unroll 1
            for (int i = tid; i < (int)cnt; i += 2 * kThreads) {
                const bool has_j = i + kThreads < (int)cnt;
                const int j = has_j ? i + kThreads : i;
                float xi, yi, zi, xj, yj, zj;
                load_pt(i, xi, yi, zi);
                load_pt(j, xj, yj, zj);  // both loads before any store: the tile is updated in place
                DistResult ri, rj;
                const bool ok_i = dist_fast<TEX, kSkipB>(L, fview, atlas, S.fast.winners, to_coxa_frame(L, xi, yi, zi), &ri);
                const bool ok_j = dist_fast<TEX, kSkipB>(L, fview, atlas, S.fast.winners, to_coxa_frame(L, xj, yj, zj), &rj);
                if (ok_i) store_pt(i, ri);
                if (ok_j & has_j) store_pt(j, rj);
                if (!ok_i) park(i);
                if (!ok_j & has_j) park(j);
            }
        } else {
#pragma unroll 1
            for (int i = tid; i < (int)cnt; i += kThreads)
                compute_point<MODE, SOA, GENERIC>(L, S.table, in, vec, flag, i, kTile);
        }

        bulk::fence_proxy_async();  // results written through the generic proxy -> bulk engine
        if (tid == 0) {
            // FAST: every committed store has COMPLETED (parked points of those tiles may be redone);
            // otherwise it only has to have drained its shared-memory buffer
            if constexpr (kRedo) {
                bulk::wait_group<0>();
                S.fast.qcnt[rot == 2 ? 0 : rot + 1] = 0;  // next iteration's counter (last read two barriers ago)
            } else {
                bulk::wait_group_read<0>();
            }
        }
        __syncthreads();
        if (tid == 0) {
            const size_t first = tile * kTile;
            if (kVec) {
                if (SOA) {
                    bulk::store(out_x + first, vec, cnt * 4);
                    bulk::store(out_y + first, vec + kTile, cnt * 4);
                    bulk::store(out_z + first, vec + 2 * kTile, cnt * 4);
                } else {
                    bulk::store(out_x + 3 * first, vec, cnt * 12);
                }
            }
            if (out_flag) bulk::store(out_flag + first, flag, cnt);
            bulk::commit_group();
            // the buffer two tiles ahead is the one whose store (issued an iteration ago) thread 0
            // has just waited for; with two buffers (reach-only) it is the one just consumed
            const size_t next = tile + (size_t)kPrefetch * gridDim.x;
            if (next < n_tiles) issue_load(next, (int)((it + kPrefetch) % kStages));
        }
        rot = rot == 2 ? 0 : rot + 1;
    }
    if (tid == 0) bulk::wait_group<0>();
    if constexpr (kRedo) {
        // drain the ring: every store has completed, every entry is final after this barrier
        __syncthreads();
        const uint32_t total = q_base + (it ? S.fast.qcnt[rot == 0 ? 2 : rot - 1] : 0u);
#pragma unroll 1
        for (; q_head < total; q_head += kThreads)
            if (q_head + tid < total) redo(S.fast.queue[(q_head + tid) % kQueueCap], it);
    }

    // the last n % 16 points bypass the bulk engine (sizes must be multiples of 16 B)
    if (blockIdx.x == 0) {
        const size_t i = n_bulk + tid;
        if (i < n) {
            float xyz[3], v[3];
            uint8_t f;
            if (SOA) {
                xyz[0] = in_x[i], xyz[1] = in_y[i], xyz[2] = in_z[i];
            } else {
                xyz[0] = in_x[3 * i], xyz[1] = in_x[3 * i + 1], xyz[2] = in_x[3 * i + 2];
            }
            compute_point<MODE, false, GENERIC>(L, S.table, xyz, v, &f, 0, 0);
            if (kVec) {
                if (SOA) {
                    out_x[i] = v[0], out_y[i] = v[1], out_z[i] = v[2];
                } else {
                    out_x[3 * i] = v[0], out_x[3 * i + 1] = v[1], out_x[3 * i + 2] = v[2];
                }
            }
            if (out_flag) out_flag[i] = f;
        }
    }
}

// ---- tiered distance sweep ----------------------------------------------------------------------
// Tier 0 (every point): ONE fetch from the choice volume returns a 16-bit texel — the winning coxa
// solution of the point's cube and, where the cube's whole plane rectangle carries one certified
// plane-atlas label, that label.  The point is then finished with one projection on the labelled
// circle / corner (dist_choice_label): no dependent second fetch, no round trip through shared
// memory.  All four texel fetches of a thread are in flight together.  What tier 0 cannot decide is
// parked in one of four per-CTA rings and redone later, 256 entries at a time by all threads of
// the CTA (dense warps, every lane on the same path), from / to global memory, and only after the
// bulk store of the entry's tile has completed (see one_leg_stream_kernel):
//   ring P: solution certified, no plane label for the cube -> the same ONE solution through the
//           plane atlas (dist_choice); hands on to ring A if the point's atlas cell is uncertified;
//   ring A: -> the same ONE solution with the explicit plane evaluation (dist_choice_clamp); cannot fail;
//   ring B: the cube is uncertified -> both solutions through the tables (dist_fast);
//   ring C: what ring B's redo cannot decide -> the full evaluation.
// A ring that is full refuses the push; the point then moves to ring C or is evaluated on the spot
// (correct, just divergent).  Every path applies the same operations to the winning candidate, so
// the output does not depend on which one ran.  A parked point's slot keeps the INPUT point, which
// the tile's store writes to the output and the redo later replaces: a call whose output buffer is
// its input buffer stays correct (the redo re-reads an unchanged input).
// CTA shape of the tiered sweep (its own: the other sweeps keep kThreads / kTile)
#ifndef LRM_TIER_THREADS
#define LRM_TIER_THREADS 256
#endif
#ifndef LRM_TIER_TILE
#define LRM_TIER_TILE 1024
#endif
#ifndef LRM_TIER_CTAS
#define LRM_TIER_CTAS 4
#endif
// points of a thread that share one "any valid plane point in the warp" vote: all four (finer
// groups measured no faster, and the one-point build faulted in its distance-only instantiation)
#ifndef LRM_T0_GROUP
#define LRM_T0_GROUP 4
#endif
constexpr int kTT = LRM_TIER_THREADS;  // threads per CTA
constexpr int kTL = LRM_TIER_TILE;     // points per tile
constexpr int kPer = kTL / kTT;        // points per thread and tile
static_assert(kTL % kTT == 0 && kTL % 16 == 0 && kTL <= 2048 && kPer % LRM_T0_GROUP == 0, "tier tile shape");
// Parked points come in runs (a lattice column that grazes a decision surface parks hundreds of
// consecutive points): rings P and B hold two such tiles.  Measured at 1e9 points: 512 / 512 entries
// 119.0, 1024 / 512 123.0, 1024 / 1024 125.0 Gpoints/s (an overflowing push falls to a slower tier).
#ifndef LRM_RING_A
#define LRM_RING_P 1024
#define LRM_RING_A 512
#define LRM_RING_B 1024
#define LRM_RING_C 512
#endif
constexpr int kRingP = LRM_RING_P, kRingA = LRM_RING_A, kRingB = LRM_RING_B, kRingC = LRM_RING_C;  // entries (powers of two)
// ring entry: iteration (17 or 16 bits) | cube byte bits 0-4 | index in tile (10 or 11 bits)
constexpr int kEntryIdxBits = kTL <= 1024 ? 10 : 11;
constexpr int kEntryIterBits = 32 - 5 - kEntryIdxBits;

// Classic 16-bit texels of a thread's NPT points: the coarse texel itself, or — where that is a brick
// pointer — the fine texel of the point's fine cube (leg_math.cuh, "bricks").  One warp vote skips the
// whole thing where no lane met a brick (most of the far field); the fine loads of a thread are
// issued together.
template <int NPT>
__device__ __forceinline__ void resolve_bricks(const VolumeView& vol, const CoxaPoint* p, unsigned* w) {
    unsigned any = 0u;
#pragma unroll
    for (int k = 0; k < NPT; k++) any |= w[k];
    if (!__any_sync(0xffffffffu, (any & kVolBrick) != 0u)) return;
    const unsigned short* src[NPT];
#pragma unroll
    for (int k = 0; k < NPT; k++)
        src[k] = vol.bricks + ((size_t)brick_index(w[k]) << 6) + brick_slot(vol, w[k], p[k]);
#pragma unroll
    for (int k = 0; k < NPT; k++)
        if (w[k] & kVolBrick) w[k] = __ldg(src[k]);
}

// tile buffers of the tiered sweep: 3 (being loaded / computed / stored) or 2 (the buffer a tile was
// stored from is reloaded as soon as the store has read it: 12 KiB less shared memory per CTA, which
// the SM hands to L1)
#ifndef LRM_TIER_STAGES
#define LRM_TIER_STAGES 2  // measured at 1e9 points: 2 buffers 134.6, 3 buffers 129.6 Gpoints/s (L1: 60 KB instead of 28 KB per SM)
#endif
constexpr int kTStages = LRM_TIER_STAGES;
static_assert(kTStages == 2 || kTStages == 3, "tile buffers");
struct alignas(128) TierSmem {
    float in[kTStages][3 * kTL];  // in-place tiles
    uint8_t flag[2][kTL];
    alignas(16) SectorTable table;
    alignas(16) WinnerTable winners;
    alignas(16) YawPair ypair[kYawPairs];
    uint32_t ring_p[kRingP], ring_a[kRingA], ring_b[kRingB], ring_c[kRingC];
    unsigned cnt[4][3];  // [ring P, A, B, C][it % 3]: pushes attempted in that iteration
    alignas(8) uint64_t full[kTStages];
};

// Per-thread (uniform) bookkeeping of one ring: accepted entries before this iteration (base),
// before the previous one (elig: their tiles' stores have completed), redone so far (head), and
// the room the ring had when this iteration's pushes began.
template <int CAP>
struct Ring {
    static_assert((CAP & (CAP - 1)) == 0 && CAP >= 2 * kTT, "ring shape");
    uint32_t base = 0, elig = 0, head = 0, room = CAP;
    // top of an iteration; prev = attempts counted in the previous one.  room is taken BEFORE this
    // iteration's redo: slots freed now may still be read by slower warps (no barrier in between).
    __device__ __forceinline__ void advance(unsigned prev) {
        elig = base;
        base += prev < room ? prev : room;
        room = CAP - (base - head);
    }
    __device__ __forceinline__ bool push(uint32_t* slots, unsigned* counter, uint32_t entry) const {
        const uint32_t k = atomicAdd(counter, 1u);
        if (k >= room) return false;
        slots[(base + k) & (CAP - 1)] = entry;
        return true;
    }
};

struct RedoIo {
    const float* __restrict__ in_x;
    const float* __restrict__ in_y;
    const float* __restrict__ in_z;
    float* __restrict__ out_x;
    float* __restrict__ out_y;
    float* __restrict__ out_z;
    uint8_t* __restrict__ out_flag;
};
template <bool SOA>
__device__ __forceinline__ CoxaPoint redo_load(const LegPlan& L, const RedoIo& io, size_t g) {
    float x, y, z;
    if (SOA) {
        x = io.in_x[g], y = io.in_y[g], z = io.in_z[g];
    } else {
        x = io.in_x[3 * g], y = io.in_x[3 * g + 1], z = io.in_x[3 * g + 2];
    }
    return to_coxa_frame(L, x, y, z);
}
template <int MODE, bool SOA>
__device__ __forceinline__ void redo_store(const RedoIo& io, size_t g, const DistResult& r) {
    if (SOA) {
        io.out_x[g] = r.dx, io.out_y[g] = r.dy, io.out_z[g] = r.dz;
    } else {
        io.out_x[3 * g] = r.dx, io.out_x[3 * g + 1] = r.dy, io.out_x[3 * g + 2] = r.dz;
    }
    if (io.out_flag) io.out_flag[g] = (MODE == kModeBoth ? r.reach : r.flag) ? 1 : 0;
}
// The redo paths are deliberately NOT inlined (see redo_point).
// ring P: the chosen solution (cube bits carried by the entry) through the plane atlas; false
// (nothing written) if the point's atlas cell is uncertified
template <int MODE, bool SOA>
__device__ __noinline__ bool redo_atlas(const LegPlan& L, const YawSol* sols, unsigned cube, const AtlasView& A,
                                        const WinnerTable& W, const RedoIo& io, size_t g) {
    DistResult r;
    if (dist_choice<true>(L, sols, cube | kVolPure, A, W, redo_load<SOA>(L, io, g), &r) != 0) return false;
    redo_store<MODE, SOA>(io, g, r);
    return true;
}
// ring A: the chosen solution with the explicit plane evaluation
template <int MODE, bool SOA>
__device__ __noinline__ void redo_choice(const LegPlan& L, const SectorTable& tab, const YawSol* sols,
                                         unsigned cube, const RedoIo& io, size_t g) {
    DistResult r;
    dist_choice_clamp(L, tab, sols, cube, redo_load<SOA>(L, io, g), &r);
    redo_store<MODE, SOA>(io, g, r);
}
// ring B: dist_fast; false (nothing written) if the tables cannot decide the point
template <int MODE, bool SOA>
__device__ __noinline__ bool redo_fast(const LegPlan& L, const FastView F, const AtlasView& A,
                                       const WinnerTable& W, const RedoIo& io, size_t g) {
    DistResult r;
    if (!dist_fast<true, false, true>(L, F, A, W, redo_load<SOA>(L, io, g), &r)) return false;
    redo_store<MODE, SOA>(io, g, r);
    return true;
}
// ring C: the full evaluation
template <int MODE, bool SOA>
__device__ __noinline__ void redo_full(const LegPlan& L, const SectorTable& tab, const RedoIo& io, size_t g) {
    redo_store<MODE, SOA>(io, g, dist_coxa_frame<false>(L, tab, redo_load<SOA>(L, io, g)));
}

// full evaluation of one point of the tile in shared memory, in place (ring overflow only); the
// slot still holds the input point
template <int MODE, bool SOA>
__device__ __noinline__ void tile_point_full(const LegPlan& L, const SectorTable& tab, float* tile,
                                             uint8_t* flag, int i) {
    float x, y, z;
    if (SOA) {
        x = tile[i], y = tile[kTL + i], z = tile[2 * kTL + i];
    } else {
        x = tile[3 * i], y = tile[3 * i + 1], z = tile[3 * i + 2];
    }
    const DistResult r = dist_coxa_frame<false>(L, tab, to_coxa_frame(L, x, y, z));
    if (SOA) {
        tile[i] = r.dx, tile[kTL + i] = r.dy, tile[2 * kTL + i] = r.dz;
    } else {
        tile[3 * i] = r.dx, tile[3 * i + 1] = r.dy, tile[3 * i + 2] = r.dz;
    }
    flag[i] = (MODE == kModeBoth ? r.reach : r.flag) ? 1 : 0;
}

// both solutions through the tables for one point of the tile in shared memory, in place (a warp
// whose points ALL lie in uncertified cubes — a sweep inside a decision surface, e.g. the
// reference's own y = 0 benchmark slice — is dense as it stands: no ring needed); false (slot
// untouched) if the tables cannot decide the point
template <int MODE, bool SOA>
__device__ __noinline__ bool tile_point_fast(const LegPlan& L, const FastView F, const AtlasView& A,
                                             const WinnerTable& W, float* tile, uint8_t* flag, int i) {
    float x, y, z;
    if (SOA) {
        x = tile[i], y = tile[kTL + i], z = tile[2 * kTL + i];
    } else {
        x = tile[3 * i], y = tile[3 * i + 1], z = tile[3 * i + 2];
    }
    DistResult r;
    if (!dist_fast<true, false, true>(L, F, A, W, to_coxa_frame(L, x, y, z), &r)) return false;
    if (SOA) {
        tile[i] = r.dx, tile[kTL + i] = r.dy, tile[2 * kTL + i] = r.dz;
    } else {
        tile[3 * i] = r.dx, tile[3 * i + 1] = r.dy, tile[3 * i + 2] = r.dz;
    }
    flag[i] = (MODE == kModeBoth ? r.reach : r.flag) ? 1 : 0;
    return true;
}

template <int MODE, bool SOA>
__global__ void __launch_bounds__(kTT, LRM_TIER_CTAS)
    one_leg_tier_kernel(const __grid_constant__ LegPlan L, const __grid_constant__ FastTables FT,
                        const AtlasView atlas, const VolumeView vol, const __grid_constant__ RedoIo io,
                        size_t n, int kshift_arg, const int* __restrict__ gate, int gate_want) {
    if (gate != nullptr && *gate != gate_want) return;  // the coherence probe chose the other sweep
    const int kshift = kshift_arg & 0xff;
#ifdef LRM_ENABLE_SKELETON
    // measurement builds only: move the tiles but skip the arithmetic — the ceiling of the staging
    // pipeline itself (results are garbage)
    const bool skeleton = (kshift_arg & 0x100) != 0;
#else
    constexpr bool skeleton = false;
#endif
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& S = *reinterpret_cast<TierSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const size_t n_bulk = n & ~size_t(15);
    const uint32_t n_tiles = (uint32_t)((n_bulk + kTL - 1) / kTL);  // the launcher keeps n below 2^40
    constexpr int kStages = kTStages;

    fill_sector_table(L, &S.table, tid, kTT);
    fill_winner_table(L, &S.winners, tid, kTT);
    for (int i = tid; i < kYawPairs * (int)(sizeof(YawPair) / 4); i += kTT)
        reinterpret_cast<float*>(S.ypair)[i] = reinterpret_cast<const float*>(FT.pair)[i];
    if (tid == 0) {
        for (int k = 0; k < 12; k++) (&S.cnt[0][0])[k] = 0;
        for (int s = 0; s < kStages; s++) bulk::mbar_init(&S.full[s], 1);
        bulk::fence_barrier_init();
    }
    __syncthreads();

    auto tile_count = [&](uint32_t tile) -> uint32_t {
        return tile + 1u < n_tiles ? (uint32_t)kTL : (uint32_t)(n_bulk - (size_t)tile * kTL);
    };
    auto issue_load = [&](uint32_t tile, int stage) {
        const uint32_t cnt = tile_count(tile);
        const size_t first = (size_t)tile * kTL;
        if (SOA) {
            bulk::mbar_expect_tx(&S.full[stage], 3 * cnt * 4);
            bulk::load(&S.in[stage][0], io.in_x + first, cnt * 4, &S.full[stage]);
            bulk::load(&S.in[stage][kTL], io.in_y + first, cnt * 4, &S.full[stage]);
            bulk::load(&S.in[stage][2 * kTL], io.in_z + first, cnt * 4, &S.full[stage]);
        } else {
            bulk::mbar_expect_tx(&S.full[stage], cnt * 12);
            bulk::load(&S.in[stage][0], io.in_x + 3 * first, cnt * 12, &S.full[stage]);
        }
    };
    // Tiles are dealt to the CTAs in chunks of 2^kshift consecutive tiles: neighbouring tiles of a
    // lattice sweep are neighbouring columns, whose points fall into the same cubes, so a CTA's
    // texture fetches keep hitting lines it has just brought in.
    auto tile_of = [&](uint32_t iter) -> uint32_t {
        return (((iter >> kshift) * gridDim.x + blockIdx.x) << kshift) + (iter & ((1u << kshift) - 1u));
    };
    if (tid == 0) {
        for (int s = 0; s < kPrefetch; s++) {
            const uint32_t tile = tile_of((uint32_t)s);
            if (tile < n_tiles) issue_load(tile, s);
        }
    }

    // the bin codes stay in the constant bank: only ring B's redo reads them
    const FastView fview{S.ypair, FT.code, FT.combo, FT.ncombo};
    const YawSol* sols = reinterpret_cast<const YawSol*>(S.ypair);
    uint32_t it = 0;
    Ring<kRingP> rp;
    Ring<kRingA> ra;
    Ring<kRingB> rb;
    Ring<kRingC> rc;
    int rot = 0;           // it % 3: also the tile buffer of this iteration
    uint32_t parity = 0;   // (it / 3) & 1: phase of that buffer's mbarrier

    constexpr uint32_t kIterMask = (1u << kEntryIterBits) - 1u;
    auto global_index = [&](uint32_t entry) -> size_t {
        const uint32_t age = (it - (entry >> (5 + kEntryIdxBits))) & kIterMask;
        return (size_t)tile_of(it - age) * kTL + (entry & ((1u << kEntryIdxBits) - 1u));
    };
    auto do_c = [&](uint32_t entry) { redo_full<MODE, SOA>(L, S.table, io, global_index(entry)); };
    auto do_b = [&](uint32_t entry) {
        if (!redo_fast<MODE, SOA>(L, fview, atlas, S.winners, io, global_index(entry)))
            if (!rc.push(S.ring_c, &S.cnt[3][rot], entry)) do_c(entry);
    };
    auto do_a = [&](uint32_t entry) {
        redo_choice<MODE, SOA>(L, S.table, sols, (entry >> kEntryIdxBits) & 31u, io, global_index(entry));
    };
    auto do_p = [&](uint32_t entry) {
        if (!redo_atlas<MODE, SOA>(L, sols, (entry >> kEntryIdxBits) & 31u, atlas, S.winners, io, global_index(entry)))
            if (!ra.push(S.ring_a, &S.cnt[1][rot], entry)) do_a(entry);
    };

    for (uint32_t tile = tile_of(0); tile < n_tiles; tile = tile_of(++it)) {
        const uint32_t cnt = tile_count(tile);
        // buffer and mbarrier phase of this tile: it % 3 and (it / 3) & 1, or it & 1 and (it >> 1) & 1
        const int buf = kStages == 3 ? rot : (int)(it & 1u);
        bulk::mbar_wait(&S.full[buf], kStages == 3 ? parity : ((it >> 1) & 1u));
        float* in = S.in[buf];
        uint8_t* flag = S.flag[it & 1];

        // counters of the previous iteration are final since its tile barrier
        const int rot_prev = rot == 0 ? 2 : rot - 1;
        rp.advance(S.cnt[0][rot_prev]);
        ra.advance(S.cnt[1][rot_prev]);
        rb.advance(S.cnt[2][rot_prev]);
        rc.advance(S.cnt[3][rot_prev]);
#pragma unroll 1
        while (rc.elig - rc.head >= (uint32_t)kTT) {
            do_c(S.ring_c[(rc.head + tid) & (kRingC - 1)]);
            rc.head += kTT;
        }
#pragma unroll 1
        while (rb.elig - rb.head >= (uint32_t)kTT) {
            do_b(S.ring_b[(rb.head + tid) & (kRingB - 1)]);
            rb.head += kTT;
        }
#pragma unroll 1
        while (ra.elig - ra.head >= (uint32_t)kTT) {
            do_a(S.ring_a[(ra.head + tid) & (kRingA - 1)]);
            ra.head += kTT;
        }
#pragma unroll 1
        while (rp.elig - rp.head >= (uint32_t)kTT) {
            do_p(S.ring_p[(rp.head + tid) & (kRingP - 1)]);
            rp.head += kTT;
        }

        auto load_pt = [&](int i, float& x, float& y, float& z) {
            if (SOA) {
                x = in[i], y = in[kTL + i], z = in[2 * kTL + i];
            } else {
                x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
            }
        };
        // a parked point leaves its slot alone (the input point): the redo rewrites it in global
        // memory after the tile's store
        auto store_pt = [&](int i, const DistResult& r) {
            if (SOA) {
                in[i] = r.dx, in[kTL + i] = r.dy, in[2 * kTL + i] = r.dz;
            } else {
                in[3 * i] = r.dx, in[3 * i + 1] = r.dy, in[3 * i + 2] = r.dz;
            }
            flag[i] = (MODE == kModeBoth ? r.reach : r.flag) ? 1 : 0;
        };
        auto park = [&](int i, int why, unsigned word) {
            const uint32_t entry = (it << (5 + kEntryIdxBits)) | ((word & 31u) << kEntryIdxBits) | (uint32_t)i;
            // why: 3 -> ring P, 1 -> ring B, 2 -> ring C (the tables have already failed)
            if (why == 3 && rp.push(S.ring_p, &S.cnt[0][rot], entry)) return;
            if (why == 3 && ra.push(S.ring_a, &S.cnt[1][rot], entry)) return;  // P full: the explicit plane evaluation
            if (why == 1 && rb.push(S.ring_b, &S.cnt[2][rot], entry)) return;
            if (rc.push(S.ring_c, &S.cnt[3][rot], entry)) return;
            tile_point_full<MODE, SOA>(L, S.table, in, flag, i);
        };
        auto fetch = [&](const CoxaPoint& p) -> unsigned {
            return tex3D<unsigned>(vol.tex, fmaf(p.x, vol.inv_cell, vol.o), fmaf(p.y, vol.inv_cell, vol.oy),
                                   fmaf(p.z, vol.inv_cell, vol.o));
        };
        if (skeleton) {
        } else if (cnt == (uint32_t)kTL) {
            // every texel of the thread is requested before the first one is consumed
            CoxaPoint p[kPer];
            unsigned w[kPer];
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                float x, y, z;
                load_pt(tid + k * kTT, x, y, z);
                p[k] = to_coxa_frame(L, x, y, z);
                w[k] = fetch(p[k]);
            }
            resolve_bricks<kPer>(vol, p, w);
#pragma unroll
            for (int g = 0; g < kPer; g += LRM_T0_GROUP) {
                // the limit-plane rule needs a valid plane point (bit 6 of the label): skipped when
                // no lane of the warp has one in this group (full tiles: the warp is converged here)
                unsigned any = 0u;
#pragma unroll
                for (int k = g; k < g + LRM_T0_GROUP; k++) any |= w[k];
                DistResult r[LRM_T0_GROUP];
                int st[LRM_T0_GROUP];
                if (__any_sync(0xffffffffu, (any & 0x4000u) != 0u)) {
#pragma unroll
                    for (int k = 0; k < LRM_T0_GROUP; k++)
                        st[k] = dist_choice_label<true>(L, sols, w[g + k], S.winners, p[g + k], &r[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < LRM_T0_GROUP; k++)
                        st[k] = dist_choice_label<false>(L, sols, w[g + k], S.winners, p[g + k], &r[k]);
                }
#pragma unroll
                for (int k = 0; k < LRM_T0_GROUP; k++)
                    if (st[k] == 0) store_pt(tid + (g + k) * kTT, r[k]);
                // a warp that is uncertified as a whole (a sweep inside a decision surface) is dense
                // already: the table path on the spot instead of 128 pushes into ring B, which such a
                // sweep would only overflow.  One vote per tile.
                bool all_b = true;
#pragma unroll
                for (int k = 0; k < LRM_T0_GROUP; k++) all_b = all_b && st[k] == 1;
                if (__all_sync(0xffffffffu, all_b)) {
#pragma unroll 1
                    for (int k = 0; k < LRM_T0_GROUP; k++) {
                        const int i = tid + (g + k) * kTT;
                        if (!tile_point_fast<MODE, SOA>(L, fview, atlas, S.winners, in, flag, i)) park(i, 2, 0u);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < LRM_T0_GROUP; k++)
                        if (st[k] != 0) park(tid + (g + k) * kTT, st[k], w[g + k]);
                }
            }
        } else {
#pragma unroll 1
            for (int i = tid; i < (int)cnt; i += kTT) {
                float x, y, z;
                load_pt(i, x, y, z);
                const CoxaPoint p = to_coxa_frame(L, x, y, z);
                unsigned w = fetch(p);
                if (w & kVolBrick) w = __ldg(vol.bricks + ((size_t)brick_index(w) << 6) + brick_slot(vol, w, p));
                DistResult r;
                const int st = dist_choice_label<true>(L, sols, w, S.winners, p, &r);
                if (st == 0) store_pt(i, r);
                else park(i, st, w);
            }
        }

        bulk::fence_proxy_async();
        if (tid == 0) {
            bulk::wait_group<0>();  // every committed store has COMPLETED: parked points of those tiles may be redone
            const int rot_next = rot == 2 ? 0 : rot + 1;
            S.cnt[0][rot_next] = 0, S.cnt[1][rot_next] = 0, S.cnt[2][rot_next] = 0, S.cnt[3][rot_next] = 0;
        }
        __syncthreads();
        if (tid == 0) {
            const size_t first = (size_t)tile * kTL;
            if (SOA) {
                bulk::store(io.out_x + first, in, cnt * 4);
                bulk::store(io.out_y + first, in + kTL, cnt * 4);
                bulk::store(io.out_z + first, in + 2 * kTL, cnt * 4);
            } else {
                bulk::store(io.out_x + 3 * first, in, cnt * 12);
            }
            if (io.out_flag) bulk::store(io.out_flag + first, flag, cnt);
            bulk::commit_group();
            static_assert(kPrefetch == 2, "two tiles ahead");
            const uint32_t next = tile_of(it + kPrefetch);
            if (kStages == 3) {
                // the buffer two tiles ahead is the previous tile's, whose store has completed
                if (next < n_tiles) issue_load(next, rot == 0 ? 2 : rot - 1);
            } else if (next < n_tiles) {
                // ... is this tile's own: reload it as soon as the store just issued has read it
                bulk::wait_group_read<0>();
                issue_load(next, buf);
            }
        }
        rot = rot == 2 ? 0 : rot + 1;
        parity ^= (rot == 0) ? 1u : 0u;
    }
    if (tid == 0) bulk::wait_group<0>();
    __syncthreads();  // every store has completed, every counter is final
    {
        const int rot_prev = rot == 0 ? 2 : rot - 1;
        if (it) {
            rp.advance(S.cnt[0][rot_prev]);
            ra.advance(S.cnt[1][rot_prev]);
            rb.advance(S.cnt[2][rot_prev]);
            rc.advance(S.cnt[3][rot_prev]);
        }
        // rings A and C first: they are final.  Then P and B, whose hand-overs go to cnt[1][rot] /
        // cnt[3][rot] (zero so far: nothing was pushed in an iteration that did not run)
#pragma unroll 1
        for (; ra.head < ra.base; ra.head += kTT)
            if (ra.head + tid < ra.base) do_a(S.ring_a[(ra.head + tid) & (kRingA - 1)]);
#pragma unroll 1
        for (; rc.head < rc.base; rc.head += kTT)
            if (rc.head + tid < rc.base) do_c(S.ring_c[(rc.head + tid) & (kRingC - 1)]);
        ra.head = ra.base, ra.room = kRingA;
        rc.head = rc.base, rc.room = kRingC;
        __syncthreads();
        // each drain round of P / B hands on at most kTT entries: rings A / C (>= 2 kTT) take them,
        // and are emptied again before the next round
#pragma unroll 1
        for (; rp.head < rp.base; rp.head += kTT) {
            if (rp.head + tid < rp.base) do_p(S.ring_p[(rp.head + tid) & (kRingP - 1)]);
            __syncthreads();
            const unsigned late = S.cnt[1][rot];
            __syncthreads();
            if (tid == 0) S.cnt[1][rot] = 0;
            if (tid < (int)late) do_a(S.ring_a[(ra.base + tid) & (kRingA - 1)]);
            __syncthreads();
        }
#pragma unroll 1
        for (; rb.head < rb.base; rb.head += kTT) {
            if (rb.head + tid < rb.base) do_b(S.ring_b[(rb.head + tid) & (kRingB - 1)]);
            __syncthreads();
            const unsigned late = S.cnt[3][rot];
            __syncthreads();
            if (tid == 0) S.cnt[3][rot] = 0;
            if (tid < (int)late) do_c(S.ring_c[(rc.base + tid) & (kRingC - 1)]);
            __syncthreads();
        }
    }

    // the last n % 16 points bypass the bulk engine
    if (blockIdx.x == 0) {
        const size_t i = n_bulk + tid;
        if (i < n) redo_full<MODE, SOA>(L, S.table, io, i);
    }
}

// ---- warp-autonomous tiered sweep ---------------------------------------------------------------
// The same tiers as one_leg_tier_kernel, organised per WARP instead of per CTA: no shared-memory
// tiles, no bulk copies, no CTA barrier, no wait for a store to complete.
//   * a warp owns tiles of 128 consecutive points; a lane owns FOUR CONSECUTIVE points of the tile:
//     48 contiguous bytes, read with three 16-byte loads and written back with three 16-byte
//     stores (AoS; SoA: one 16-byte load / store per plane), flags as one 32-bit word — every
//     sector of the streams is touched by exactly one warp, once;
//   * tier 0 as above (one texel fetch per point, all four in flight);
//   * what tier 0 cannot decide goes to warp-private rings in shared memory (index only); the
//     lane writes the INPUT back for such a point, and once a ring holds a warp's worth of entries
//     they are redone from / to global memory by the same warp, 32 at a time — after the
//     tile's own stores in program order (__syncwarp orders the two writes), so there is nothing to
//     wait for and a call whose output buffer is its input buffer stays correct;
//   * rings: P (solution certified, cube without plane label -> plane atlas; hands on to A),
//     A (-> explicit plane evaluation), B (cube uncertified -> dist_fast; hands on to C),
//     C (-> full evaluation).
#ifndef LRM_WARP_CTAS
#define LRM_WARP_CTAS 4
#endif
constexpr int kWT = 128;     // points per warp-tile
constexpr int kWW = 8;       // warps per CTA
constexpr int kWCapBig = 256, kWCapSmall = 64;  // ring capacities (powers of two)
struct WarpRings {
    uint32_t p[kWCapBig], b[kWCapBig], a[kWCapSmall], c[kWCapSmall];
};
struct alignas(128) WarpSmem {
    alignas(16) SectorTable table;
    alignas(16) WinnerTable winners;
    alignas(16) YawPair ypair[kYawPairs];
    WarpRings ring[kWW];
};
// ring entry: warp iteration (20 bits) | cube byte bits 0-4 | point in tile (7 bits)
struct WarpGeom {
    uint32_t w, nw;  // this warp, all warps
    int kshift;
    __device__ __forceinline__ uint32_t tile_of(uint32_t iter) const {
        return (((iter >> kshift) * nw + w) << kshift) + (iter & ((1u << kshift) - 1u));
    }
    __device__ __forceinline__ size_t global_index(uint32_t entry) const {
        return (size_t)tile_of(entry >> 12) * kWT + (entry & 127u);
    }
};

template <int MODE, bool SOA>
__global__ void __launch_bounds__(kWW * 32, LRM_WARP_CTAS)
    one_leg_warp_kernel(const __grid_constant__ LegPlan L, const __grid_constant__ FastTables FT,
                        const AtlasView atlas, const VolumeView vol, const __grid_constant__ RedoIo io,
                        size_t n, int kshift, const int* __restrict__ gate, int gate_want) {
    if (gate != nullptr && *gate != gate_want) return;  // the coherence probe chose the other sweep
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto& S = *reinterpret_cast<WarpSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    fill_sector_table(L, &S.table, tid, kWW * 32);
    fill_winner_table(L, &S.winners, tid, kWW * 32);
    for (int i = tid; i < kYawPairs * (int)(sizeof(YawPair) / 4); i += kWW * 32)
        reinterpret_cast<float*>(S.ypair)[i] = reinterpret_cast<const float*>(FT.pair)[i];
    __syncthreads();

    const uint32_t n_tiles = (uint32_t)(n / kWT);  // full tiles; the launcher keeps n below 2^38
    const WarpGeom G{(uint32_t)(blockIdx.x * kWW + warp), (uint32_t)(gridDim.x * kWW), kshift};
    WarpRings& R = S.ring[warp];
    const FastView fview{S.ypair, FT.code, FT.combo, FT.ncombo};
    const YawSol* sols = reinterpret_cast<const YawSol*>(S.ypair);
    const unsigned lt = (1u << lane) - 1u;
    uint32_t hp = 0, tp = 0, ha = 0, ta = 0, hb = 0, tb = 0, hc = 0, tc = 0;  // ring heads / tails (warp-uniform)

    // one batch of up to 32 entries of a ring; `cnt` lanes take part
    auto batch_c = [&](uint32_t cnt) {
        if ((uint32_t)lane < cnt) redo_full<MODE, SOA>(L, S.table, io, G.global_index(R.c[(hc + lane) & (kWCapSmall - 1)]));
        hc += cnt;
        __syncwarp();
    };
    auto batch_a = [&](uint32_t cnt) {
        if ((uint32_t)lane < cnt) {
            const uint32_t e = R.a[(ha + lane) & (kWCapSmall - 1)];
            redo_choice<MODE, SOA>(L, S.table, sols, (e >> 7) & 31u, io, G.global_index(e));
        }
        ha += cnt;
        __syncwarp();
    };
    auto batch_b = [&](uint32_t cnt) {
        uint32_t e = 0;
        bool fail = false;
        if ((uint32_t)lane < cnt) {
            e = R.b[(hb + lane) & (kWCapBig - 1)];
            fail = !redo_fast<MODE, SOA>(L, fview, atlas, S.winners, io, G.global_index(e));
        }
        hb += cnt;
        const unsigned m = __ballot_sync(0xffffffffu, fail);
        if (m) {
            if (tc - hc + __popc(m) > (uint32_t)kWCapSmall) batch_c(tc - hc < 32u ? tc - hc : 32u);
            if (fail) R.c[(tc + __popc(m & lt)) & (kWCapSmall - 1)] = e;
            tc += __popc(m);
        }
        __syncwarp();
    };
    auto batch_p = [&](uint32_t cnt) {
        uint32_t e = 0;
        bool fail = false;
        if ((uint32_t)lane < cnt) {
            e = R.p[(hp + lane) & (kWCapBig - 1)];
            fail = !redo_atlas<MODE, SOA>(L, sols, (e >> 7) & 31u, atlas, S.winners, io, G.global_index(e));
        }
        hp += cnt;
        const unsigned m = __ballot_sync(0xffffffffu, fail);
        if (m) {
            if (ta - ha + __popc(m) > (uint32_t)kWCapSmall) batch_a(ta - ha < 32u ? ta - ha : 32u);
            if (fail) R.a[(ta + __popc(m & lt)) & (kWCapSmall - 1)] = e;
            ta += __popc(m);
        }
        __syncwarp();
    };

    for (uint32_t it = 0;; ++it) {
        const uint32_t tile = G.tile_of(it);
        if (tile >= n_tiles) break;
        const size_t base = (size_t)tile * kWT;
        // the lane's four consecutive points
        float x[4], y[4], z[4];
        if (SOA) {
            const float4 vx = __ldcs(reinterpret_cast<const float4*>(io.in_x + base) + lane);
            const float4 vy = __ldcs(reinterpret_cast<const float4*>(io.in_y + base) + lane);
            const float4 vz = __ldcs(reinterpret_cast<const float4*>(io.in_z + base) + lane);
            x[0] = vx.x, x[1] = vx.y, x[2] = vx.z, x[3] = vx.w;
            y[0] = vy.x, y[1] = vy.y, y[2] = vy.z, y[3] = vy.w;
            z[0] = vz.x, z[1] = vz.y, z[2] = vz.z, z[3] = vz.w;
        } else {
            const float4* src = reinterpret_cast<const float4*>(io.in_x + 3 * base) + 3 * lane;
            const float4 v0 = __ldcs(src), v1 = __ldcs(src + 1), v2 = __ldcs(src + 2);
            x[0] = v0.x, y[0] = v0.y, z[0] = v0.z, x[1] = v0.w, y[1] = v1.x, z[1] = v1.y;
            x[2] = v1.z, y[2] = v1.w, z[2] = v2.x, x[3] = v2.y, y[3] = v2.z, z[3] = v2.w;
        }
        CoxaPoint p[4];
        unsigned wd[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            p[k] = to_coxa_frame(L, x[k], y[k], z[k]);
            wd[k] = tex3D<unsigned>(vol.tex, fmaf(p[k].x, vol.inv_cell, vol.o), fmaf(p[k].y, vol.inv_cell, vol.oy),
                                    fmaf(p[k].z, vol.inv_cell, vol.o));
        }
        resolve_bricks<4>(vol, p, wd);
        DistResult r[4];
        int st[4];
        // the limit-plane rule needs a valid plane point (bit 6 of the label): skipped when no lane
        // of the warp has one in this tile
        if (__any_sync(0xffffffffu, ((wd[0] | wd[1] | wd[2] | wd[3]) & 0x4000u) != 0u)) {
#pragma unroll
            for (int k = 0; k < 4; k++) st[k] = dist_choice_label<true>(L, sols, wd[k], S.winners, p[k], &r[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) st[k] = dist_choice_label<false>(L, sols, wd[k], S.winners, p[k], &r[k]);
        }
        // a parked point gets its input written back (the redo replaces it)
        unsigned fl = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (st[k] == 0) {
                x[k] = r[k].dx, y[k] = r[k].dy, z[k] = r[k].dz;
                fl |= ((MODE == kModeBoth ? r[k].reach : r[k].flag) ? 1u : 0u) << (8 * k);
            }
        }
        if (SOA) {
            __stcs(reinterpret_cast<float4*>(io.out_x + base) + lane, make_float4(x[0], x[1], x[2], x[3]));
            __stcs(reinterpret_cast<float4*>(io.out_y + base) + lane, make_float4(y[0], y[1], y[2], y[3]));
            __stcs(reinterpret_cast<float4*>(io.out_z + base) + lane, make_float4(z[0], z[1], z[2], z[3]));
        } else {
            float4* dst = reinterpret_cast<float4*>(io.out_x + 3 * base) + 3 * lane;
            __stcs(dst, make_float4(x[0], y[0], z[0], x[1]));
            __stcs(dst + 1, make_float4(y[1], z[1], x[2], y[2]));
            __stcs(dst + 2, make_float4(z[2], x[3], y[3], z[3]));
        }
        if (io.out_flag) __stcs(reinterpret_cast<unsigned*>(io.out_flag + base) + lane, fl);
        if (__any_sync(0xffffffffu, (st[0] | st[1] | st[2] | st[3]) != 0)) {
            __syncwarp();  // the tile's stores precede the redo's stores to the same addresses
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t e = (it << 12) | ((wd[k] & 31u) << 7) | (uint32_t)(4 * lane + k);
                const unsigned mp = __ballot_sync(0xffffffffu, st[k] == 3);
                if (st[k] == 3) R.p[(tp + __popc(mp & lt)) & (kWCapBig - 1)] = e;
                tp += __popc(mp);
                const unsigned mb = __ballot_sync(0xffffffffu, st[k] == 1);
                if (st[k] == 1) R.b[(tb + __popc(mb & lt)) & (kWCapBig - 1)] = e;
                tb += __popc(mb);
            }
            __syncwarp();
#pragma unroll 1
            while (tp - hp >= 32u) batch_p(32u);
#pragma unroll 1
            while (ta - ha >= 32u) batch_a(32u);
#pragma unroll 1
            while (tb - hb >= 32u) batch_b(32u);
#pragma unroll 1
            while (tc - hc >= 32u) batch_c(32u);
        }
    }
    // drain: P before A, B before C (the hand-overs)
#pragma unroll 1
    while (tp != hp) batch_p(tp - hp < 32u ? tp - hp : 32u);
#pragma unroll 1
    while (ta != ha) batch_a(ta - ha < 32u ? ta - ha : 32u);
#pragma unroll 1
    while (tb != hb) batch_b(tb - hb < 32u ? tb - hb : 32u);
#pragma unroll 1
    while (tc != hc) batch_c(tc - hc < 32u ? tc - hc : 32u);

    // the last n % 128 points
    if (blockIdx.x == 0) {
        const size_t i = (size_t)n_tiles * kWT + tid;
        if (i < n) redo_full<MODE, SOA>(L, S.table, io, i);
    }
}

// Same math with plain per-thread global accesses: used when a caller's buffers are not 16-byte
// aligned (the bulk engine's requirement) — still the GPU path, just without staging.
template <int MODE, bool GENERIC>
__global__ void __launch_bounds__(kThreads)
    one_leg_plain_kernel(const __grid_constant__ LegPlan L, const float* __restrict__ xyz,
                         float* __restrict__ out_vec, uint8_t* __restrict__ out_flag, size_t n) {
    __shared__ SectorTable table;
    fill_sector_table(L, &table, threadIdx.x, kThreads);
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        float v[3];
        uint8_t f;
        compute_point<MODE, false, GENERIC>(L, table, p, v, &f, 0, 0);
        if (MODE & kModeDist) {
            out_vec[3 * i] = v[0], out_vec[3 * i + 1] = v[1], out_vec[3 * i + 2] = v[2];
        }
        if (out_flag) out_flag[i] = f;
    }
}

// forward_kine_kernel, one_leg.cu:377-414 (coxa_pitch is ignored there too)
__global__ void forward_kine_kernel_b200(const float* __restrict__ angles, lrm_leg_t leg,
                                         float* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float coxa = angles[3 * i], femur = angles[3 * i + 1], tibia = angles[3 * i + 2];
        float sh, ch, s, c;
        sincosf(coxa, &sh, &ch);
        float x = leg.body + ch * leg.coxa_length, y = sh * leg.coxa_length, z = 0.f;
        sincosf(femur, &s, &c);
        float horiz = c * leg.femur_length, vert = s * leg.femur_length;
        x += ch * horiz, y += sh * horiz, z += vert;
        sincosf(tibia + femur, &s, &c);
        horiz = c * leg.tibia_length, vert = s * leg.tibia_length;
        x += ch * horiz, y += sh * horiz, z += vert;
        out[3 * i] = x, out[3 * i + 1] = y, out[3 * i + 2] = z;
    }
}

// apply_recurs / recursive_kernel / fillOutKernel (cross_compiled.cu:82-139,
// one_leg_global.cu:168-251, octree_util.cu:9-26): an adaptive octree of the single-leg distance
// field (root +-5000 mm, axes stop splitting below 100 mm) whose leaves paint (depth, 0, 0) on the
// query points they contain.  The reference materialises the tree with dynamic parallelism and
// runs one full pass over ALL query points per leaf.  The boxes tile space ((-h, h] per axis), so
// here every point walks down its own branch without any tree: pick the child holding the point,
// evaluate the distance field at its centre, stop when the reachability edge cannot cross the
// child (|d| >= |half extents|), nothing was split, or the depth limit is hit.
template <bool GENERIC>
__global__ void __launch_bounds__(kThreads)
    recurs_kernel(const __grid_constant__ LegPlan L, const float* __restrict__ xyz,
                  float* __restrict__ out, size_t n, int max_depth) {
    __shared__ SectorTable table;
    fill_sector_table(L, &table, threadIdx.x, kThreads);
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        float c[3] = {0.f, 0.f, 0.f}, h[3] = {5000.f, 5000.f, 5000.f};  // BoxCenter / BoxSize, settings.h:25-26
        bool inside = true;
        for (int q = 0; q < 3; q++) inside = inside && (h[q] >= p[q] - c[q]) && (-h[q] < p[q] - c[q]);
        if (!inside) continue;  // the reference paints nothing there either
        int depth = 0;
        while (true) {
            int nsplit = 0;
            for (int q = 0; q < 3; q++) {
                if (fabsf(h[q]) < 100.f) continue;  // MIN_BOX: this axis is not split any more
                const float nh = h[q] / 2.f, mv = h[q] - nh;
                c[q] += (p[q] - c[q] > 0.f) ? mv : -mv;
                h[q] = nh;
                nsplit++;
            }
            const DistResult d = dist_coxa_frame<GENERIC>(L, table, to_coxa_frame(L, c[0], c[1], c[2]));
            const bool edge_in_box = norm3df(d.dx, d.dy, d.dz) < norm3df(h[0], h[1], h[2]);
            if (edge_in_box && nsplit > 0 && depth < max_depth) {
                depth++;
            } else {
                break;
            }
        }
        out[3 * i] = (float)depth, out[3 * i + 1] = 0.f, out[3 * i + 2] = 0.f;
    }
}

// generate3DGrid (bench.cpp:30-50) on the device: x-major, z fastest
__global__ void lattice_kernel(float* __restrict__ out, float3 lo, float3 step, uint32_t ny,
                               uint32_t nz, size_t first, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const size_t i = first + k;
        const uint32_t iz = (uint32_t)(i % nz);
        const size_t t = i / nz;
        const uint32_t iy = (uint32_t)(t % ny);
        const uint32_t ix = (uint32_t)(t / ny);
        // separately rounded multiply and add: a host loop with the same two operations matches
        out[3 * k] = __fadd_rn(lo.x, __fmul_rn((float)ix, step.x));
        out[3 * k + 1] = __fadd_rn(lo.y, __fmul_rn((float)iy, step.y));
        out[3 * k + 2] = __fadd_rn(lo.z, __fmul_rn((float)iz, step.z));
    }
}

int g_sm_count = 0;
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <int MODE, bool SOA, bool GENERIC, bool FAST, bool TEX>
cudaError_t launch_stream_impl(const LegPlan& plan, const FastTables& ft, const AtlasView& atlas,
                               const float* ix, const float* iy, const float* iz, float* ox, float* oy,
                               float* oz, uint8_t* flag, size_t n, cudaStream_t stream,
                               const int* gate = nullptr, int gate_want = 0, const VolumeView* vol = nullptr) {
    auto kernel = one_leg_stream_kernel<MODE, SOA, GENERIC, FAST, TEX>;
    constexpr size_t smem = sizeof(StreamSmem<MODE, FAST>);
    // per-device: the attribute belongs to the device's copy of the function
    static int ctas_per_sm_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& ctas_per_sm = ctas_per_sm_dev[dev & 63];
    if (ctas_per_sm == 0) {
        cudaError_t e =
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem);
        if (e != cudaSuccess) return e;
        ctas_per_sm = occ < 1 ? 1 : occ;
    }
    const size_t tiles = ((n & ~size_t(15)) + kTile - 1) / kTile;
    size_t grid = (size_t)sm_count() * ctas_per_sm;
    if (tiles < grid) grid = tiles;
    if (grid == 0) grid = 1;
    VolumeView no_volume{};
    kernel<<<(unsigned)grid, kThreads, smem, stream>>>(plan, ft, atlas, ix, iy, iz, ox, oy, oz, flag, n, gate,
                                                       gate_want, vol ? *vol : no_volume);
    return cudaGetLastError();
}

// Options (lrm_set_option): log2 of the tiles per chunk of the tiered sweep, and which sweep large
// distance calls take (0 = always the two-tier sweep, 1 = always the tiered sweep, 2 = decided per
// launch by the coherence probe).
std::atomic<int> g_tier_chunk_shift{3};
std::atomic<int> g_sweep_mode{2};
#ifdef LRM_ENABLE_SKELETON
std::atomic<int> g_skeleton{0};
#endif

std::atomic<int> g_tier_kernel{0};  // 0 = CTA tiles + CTA rings (one_leg_tier_kernel, the faster one: 118 vs 110-114 Gpoints/s), 1 = warp-autonomous

template <int MODE, bool SOA>
cudaError_t launch_warp_impl(const LegPlan& plan, const FastTables& ft, const AtlasView& atlas,
                             const VolumeView& vol, const float* ix, const float* iy, const float* iz,
                             float* ox, float* oy, float* oz, uint8_t* flag, size_t n, cudaStream_t stream,
                             const int* gate, int gate_want) {
    auto kernel = one_leg_warp_kernel<MODE, SOA>;
    constexpr size_t smem = sizeof(WarpSmem);
    static int ctas_per_sm_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& ctas_per_sm = ctas_per_sm_dev[dev & 63];
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kWW * 32, smem);
        if (e != cudaSuccess) return e;
        ctas_per_sm = occ < 1 ? 1 : occ;
    }
    const size_t tiles = n / kWT;
    size_t grid = (size_t)sm_count() * ctas_per_sm;
    if ((tiles + kWW - 1) / kWW < grid) grid = (tiles + kWW - 1) / kWW;
    if (grid == 0) grid = 1;
    // chunks of up to 8 consecutive tiles (one lattice column) per warp, fewer on small sweeps
    int kshift = 0;
    const int kshift_max = g_tier_chunk_shift.load(std::memory_order_relaxed);
    while (kshift < kshift_max && (tiles >> (kshift + 1)) >= grid * kWW * 8) kshift++;
    const RedoIo io{ix, iy, iz, ox, oy, oz, flag};
    kernel<<<(unsigned)grid, kWW * 32, smem, stream>>>(plan, ft, atlas, vol, io, n, kshift, gate, gate_want);
    return cudaGetLastError();
}

template <int MODE, bool SOA>
cudaError_t launch_tier_impl(const LegPlan& plan, const FastTables& ft, const AtlasView& atlas,
                             const VolumeView& vol, const float* ix, const float* iy, const float* iz,
                             float* ox, float* oy, float* oz, uint8_t* flag, size_t n, cudaStream_t stream,
                             const int* gate, int gate_want) {
    if (g_tier_kernel.load(std::memory_order_relaxed) == 1 && n < (size_t(1) << 38))
        return launch_warp_impl<MODE, SOA>(plan, ft, atlas, vol, ix, iy, iz, ox, oy, oz, flag, n, stream, gate, gate_want);
    auto kernel = one_leg_tier_kernel<MODE, SOA>;
    constexpr size_t smem = sizeof(TierSmem);
    static int ctas_per_sm_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& ctas_per_sm = ctas_per_sm_dev[dev & 63];
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTT, smem);
        if (e != cudaSuccess) return e;
        ctas_per_sm = occ < 1 ? 1 : occ;
    }
    const size_t tiles = ((n & ~size_t(15)) + kTL - 1) / kTL;
    size_t grid = (size_t)sm_count() * ctas_per_sm;
    if (tiles < grid) grid = tiles;
    if (grid == 0) grid = 1;
    // chunks of up to 8 consecutive tiles per CTA, fewer on small sweeps (keep >= 16 chunks per CTA)
    int kshift = 0;
    const int kshift_max = g_tier_chunk_shift.load(std::memory_order_relaxed);
    while (kshift < kshift_max && (tiles >> (kshift + 1)) >= grid * 16) kshift++;
#ifdef LRM_ENABLE_SKELETON  // measurement builds only (LRM_NVCC_EXTRA=-DLRM_ENABLE_SKELETON): results are garbage
    if (g_skeleton.load()) kshift |= 0x100;
#endif
    const RedoIo io{ix, iy, iz, ox, oy, oz, flag};
    kernel<<<(unsigned)grid, kTT, smem, stream>>>(plan, ft, atlas, vol, io, n, kshift, gate, gate_want);
    return cudaGetLastError();
}

// Which sweep suits the input?  The tiered sweep adds a fetch from the 3-D choice volume (hundreds
// of MB): a win when neighbouring points of the array are neighbours in space (lattices, scan lines,
// sorted clouds: the fetches of a warp share a few sectors), a loss on shuffled clouds, where every
// fetch is its own DRAM sector.  256 pairs of consecutive points spread over the array vote; the
// verdict goes to a device word that both sweeps read first — the one that is not chosen returns
// at once.  No host synchronisation: device-pointer calls stay asynchronous.
template <bool SOA>
__global__ void coherence_probe_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                       const float* __restrict__ z, size_t n, float near_mm, int* verdict) {
    const size_t s = n < 2 ? 0 : (size_t)threadIdx.x * ((n - 2) / blockDim.x);
    const size_t t = s + 1 < n ? s + 1 : s;
    float dx, dy, dz;
    if (SOA) {
        dx = x[s] - x[t], dy = y[s] - y[t], dz = z[s] - z[t];
    } else {
        dx = x[3 * s] - x[3 * t], dy = x[3 * s + 1] - x[3 * t + 1], dz = x[3 * s + 2] - x[3 * t + 2];
    }
    const bool near = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) < near_mm * near_mm;
    const int votes = __syncthreads_count(near);
    if (threadIdx.x == 0) *verdict = (4 * votes >= 3 * (int)blockDim.x) ? 1 : 0;
}

// verdict words: a small per-device pool handed out round-robin (a word is reused 4096 launches later)
int* next_verdict_word() {
    constexpr int kWords = 4096;
    static int* pool[64] = {nullptr};
    static std::atomic<unsigned> next{0};
    int dev = 0;
    cudaGetDevice(&dev);
    int*& p = pool[dev & 63];
    if (p == nullptr && cudaMalloc((void**)&p, kWords * sizeof(int)) != cudaSuccess) {
        p = nullptr;
        (void)cudaGetLastError();
        return nullptr;
    }
    return p + (next.fetch_add(1u, std::memory_order_relaxed) % kWords);
}

// A leased set of certified tables, released (and its use recorded on the stream) when the
// launcher returns — after its last launch.
struct Tables {
    TableLease lease;
    cudaStream_t stream;
    AtlasView atlas{};
    FastTables ft{};
    explicit Tables(cudaStream_t s) : stream(s) {}
    cudaError_t acquire(const LegPlan& plan) { return acquire_tables(plan, stream, &atlas, &ft, &lease); }
    ~Tables() { release_tables(&lease, stream); }
    Tables(const Tables&) = delete;
    Tables& operator=(const Tables&) = delete;
};

// reach-only never cross-validates, so only the distance modes have a generic instantiation
template <int MODE, bool SOA>
cudaError_t launch_stream(const LegPlan& plan, const float* ix, const float* iy, const float* iz,
                          float* ox, float* oy, float* oz, uint8_t* flag, size_t n,
                          cudaStream_t stream, size_t n_call = 0) {
    AtlasView none{};
    static const FastTables no_tables{};
    const size_t n_eff = n_call > n ? n_call : n;
    const size_t n_min = g_fast_min_points.load(std::memory_order_relaxed);
    const bool big = n_eff >= n_min ||
                     (n_eff >= kCachedMinPoints && n_min == kAtlasMinPoints && !plan.generic && tables_cached(plan));
    const int vmode = g_sweep_mode.load(std::memory_order_relaxed);
    if (MODE == kModeReach) {
        // large reach-only sweeps read the valid bit of the plane atlas instead of testing circles
        if (big && !plan.generic) {
            Tables T(stream);
            cudaError_t e = T.acquire(plan);
            if (e != cudaSuccess) return e;
            // reach bits of the choice volume (built in the background on first request); as for
            // the distance sweeps, only for inputs the coherence probe finds spatially ordered
            VolumeView vol{};
            if (vmode != 0 && get_choice_volume(T.lease, stream, &vol, /*wait=*/vmode == 1) == cudaSuccess) {
                int* verdict = vmode == 2 ? next_verdict_word() : nullptr;
                if (verdict != nullptr)
                    coherence_probe_kernel<SOA><<<1, 256, 0, stream>>>(ix, iy, iz, n, 2.0f / vol.inv_cell, verdict);
                if (vmode == 1 || verdict != nullptr)
                    return launch_stream_impl<MODE, SOA, false, MODE == kModeReach, true>(
                        plan, no_tables, T.atlas, ix, iy, iz, ox, oy, oz, flag, n, stream, verdict, -1, &vol);
            }
            (void)cudaGetLastError();
            return launch_stream_impl<MODE, SOA, false, MODE == kModeReach, true>(
                plan, no_tables, T.atlas, ix, iy, iz, ox, oy, oz, flag, n, stream);
        }
        return launch_stream_impl<MODE, SOA, false, false, false>(plan, no_tables, none, ix, iy, iz, ox, oy,
                                                                  oz, flag, n, stream);
    }
    constexpr bool kDist = MODE != kModeReach;
    if (plan.generic)
        return launch_stream_impl<MODE, SOA, kDist, false, false>(plan, no_tables, none, ix, iy, iz, ox, oy,
                                                                  oz, flag, n, stream);
    // the atlas pays for itself (16 Mi probes) only on large sweeps, or once it is cached; ring
    // entries hold a 22-bit per-CTA iteration count (n / (kTile * grid) is far below that)
    if (big && n / kTile / (size_t)sm_count() < (size_t(1) << 21)) {
        Tables T(stream);
        cudaError_t e = T.acquire(plan);
        if (e != cudaSuccess) return e;
        if constexpr (kDist) {
            // ring entries of the tiered sweep hold a 17-bit per-CTA iteration count
            if (vmode != 0 && T.ft.both_unsat == 0 &&
                n / kTL / (size_t)sm_count() < (size_t(1) << (kEntryIterBits - 1)) && n < (size_t(1) << 40)) {
                VolumeView vol;
                e = get_choice_volume(T.lease, stream, &vol, /*wait=*/vmode == 1);
                int* verdict = (e == cudaSuccess && vmode == 2) ? next_verdict_word() : nullptr;
                if (e == cudaSuccess && vmode == 2 && verdict != nullptr) {
                    // "near" = within two cubes of the volume
                    coherence_probe_kernel<SOA><<<1, 256, 0, stream>>>(ix, iy, iz, n, 2.0f / vol.inv_cell, verdict);
                    e = launch_tier_impl<MODE, SOA>(plan, T.ft, T.atlas, vol, ix, iy, iz, ox, oy, oz, flag, n, stream,
                                                    verdict, 1);
                    if (e != cudaSuccess) return e;
                    return launch_stream_impl<MODE, SOA, false, kDist, kDist>(plan, T.ft, T.atlas, ix, iy, iz, ox, oy,
                                                                              oz, flag, n, stream, verdict, 0);
                }
                if (e == cudaSuccess)
                    return launch_tier_impl<MODE, SOA>(plan, T.ft, T.atlas, vol, ix, iy, iz, ox, oy, oz, flag, n,
                                                       stream, nullptr, 0);
                (void)cudaGetLastError();  // volume still being built, or no memory for it: the two-tier sweep gives the same bits
            }
        }
        return launch_stream_impl<MODE, SOA, false, kDist, kDist>(plan, T.ft, T.atlas, ix, iy, iz, ox, oy, oz,
                                                                  flag, n, stream);
    }
    return launch_stream_impl<MODE, SOA, false, false, false>(plan, no_tables, none, ix, iy, iz, ox, oy, oz,
                                                              flag, n, stream);
}

template <int MODE>
cudaError_t launch_plain(const LegPlan& plan, const float* xyz, float* out_vec, uint8_t* flag,
                         size_t n, cudaStream_t stream) {
    size_t grid = (n + kThreads - 1) / kThreads;
    const size_t cap = (size_t)sm_count() * 16;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    if (MODE != kModeReach && plan.generic)
        one_leg_plain_kernel<MODE, MODE != kModeReach>
            <<<(unsigned)grid, kThreads, 0, stream>>>(plan, xyz, out_vec, flag, n);
    else
        one_leg_plain_kernel<MODE, false>
            <<<(unsigned)grid, kThreads, 0, stream>>>(plan, xyz, out_vec, flag, n);
    return cudaGetLastError();
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

size_t set_fast_path_min_points(size_t n) { return g_fast_min_points.exchange(n); }
int set_sweep_mode(int mode) { return g_sweep_mode.exchange(mode < 0 ? 0 : (mode > 2 ? 2 : mode)); }
int set_tier_kernel(int which) { return g_tier_kernel.exchange(which ? 1 : 0); }
int set_tier_chunk_shift(int shift) { return g_tier_chunk_shift.exchange(shift < 0 ? 0 : (shift > 8 ? 8 : shift)); }
int set_skeleton(int on) {
#ifdef LRM_ENABLE_SKELETON
    return g_skeleton.exchange(on);
#else
    (void)on;
    return -1;
#endif
}

cudaError_t launch_one_leg_aos(int mode, const LegPlan& plan, const float* xyz, float* out_vec,
                               uint8_t* flag, size_t n, cudaStream_t stream, size_t n_call) {
    if (n == 0) return cudaSuccess;
    const bool bulk_ok = aligned16(xyz) && aligned16(out_vec) && aligned16(flag) &&
                         (n_call > n ? n_call : n) >= kPlainMaxPoints;
    switch (mode) {
        case kModeReach:
            return bulk_ok ? launch_stream<kModeReach, false>(plan, xyz, nullptr, nullptr, nullptr,
                                                              nullptr, nullptr, flag, n, stream, n_call)
                           : launch_plain<kModeReach>(plan, xyz, nullptr, flag, n, stream);
        case kModeDist:
            return bulk_ok ? launch_stream<kModeDist, false>(plan, xyz, nullptr, nullptr, out_vec,
                                                             nullptr, nullptr, flag, n, stream, n_call)
                           : launch_plain<kModeDist>(plan, xyz, out_vec, flag, n, stream);
        case kModeBoth:
            return bulk_ok ? launch_stream<kModeBoth, false>(plan, xyz, nullptr, nullptr, out_vec,
                                                             nullptr, nullptr, flag, n, stream, n_call)
                           : launch_plain<kModeBoth>(plan, xyz, out_vec, flag, n, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_one_leg_soa(const LegPlan& plan, const float* x, const float* y, const float* z,
                               float* dx, float* dy, float* dz, uint8_t* flag, size_t n,
                               cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (!(aligned16(x) && aligned16(y) && aligned16(z) && aligned16(dx) && aligned16(dy) &&
          aligned16(dz) && aligned16(flag)))
        return cudaErrorMisalignedAddress;
    if (dx == nullptr)
        return launch_stream<kModeReach, true>(plan, x, y, z, nullptr, nullptr, nullptr, flag, n,
                                               stream);
    return launch_stream<kModeBoth, true>(plan, x, y, z, dx, dy, dz, flag, n, stream);
}

cudaError_t launch_recurs(const LegPlan& plan, const float* xyz, float* out, size_t n, int max_depth,
                          cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    size_t grid = (n + kThreads - 1) / kThreads;
    if (grid > (size_t)sm_count() * 8) grid = (size_t)sm_count() * 8;
    if (plan.generic)
        recurs_kernel<true><<<(unsigned)grid, kThreads, 0, stream>>>(plan, xyz, out, n, max_depth);
    else
        recurs_kernel<false><<<(unsigned)grid, kThreads, 0, stream>>>(plan, xyz, out, n, max_depth);
    return cudaGetLastError();
}

cudaError_t launch_forward_kine(const float* angles, const lrm_leg_t& leg, float* out, size_t n,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    size_t grid = (n + 255) / 256;
    if (grid > (size_t)sm_count() * 16) grid = (size_t)sm_count() * 16;
    forward_kine_kernel_b200<<<(unsigned)grid, 256, 0, stream>>>(angles, leg, out, n);
    return cudaGetLastError();
}

cudaError_t launch_lattice(float* out, const float lo[3], const float step[3],
                           const uint32_t dims[3], size_t first, size_t count,
                           cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    size_t grid = (count + 255) / 256;
    if (grid > (size_t)sm_count() * 32) grid = (size_t)sm_count() * 32;
    lattice_kernel<<<(unsigned)grid, 256, 0, stream>>>(out, make_float3(lo[0], lo[1], lo[2]),
                                                       make_float3(step[0], step[1], step[2]),
                                                       dims[1], dims[2], first, count);
    return cudaGetLastError();
}

}  // namespace lrm
