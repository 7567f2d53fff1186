// positionability.cu — multi-leg body-positionability search over a point-cloud map.
//
// Replaces multi_rot_estimator / robot_full_struct (several_leg.cu:326-877) and the kernels it
// drives: in_sphere_mem_kernel (collision.cu:40-66), the 2-functor double_reduction_kernel
// (cuda_util.cuh:159-244), reach_mem_kernel (several_leg.cu:92-129) and the thrust
// rotate / partition / compact passes between them.
//
// The reference answers every "exists a map point such that ..." question by brute force: one
// block per body position, every block streams the whole map from global memory, once per
// predicate, per leg, per orientation, and both clouds are re-rotated and re-compacted by thrust
// for each of the 45 orientations.  Here the map is bucketed once into an xy cell grid
// (cell-sorted float4 points + per-cell z range, small enough to stay L2-resident), and one warp
// owns one body position for the whole search: it walks only the cells that can contain a
// witness, 32 map points per step, with __any_sync / __ballot_sync early exit per predicate, per
// leg and per orientation, entirely in registers.  Per-(orientation, leg) constants (fused
// rotation + leg frame, oriented tibia limits, circle tables) are built on the host.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "cell_grid.cuh"
#include "kernels.h"
#include "leg_math.cuh"

namespace lrm {

namespace {

#ifndef LRM_POSIT_WARPS
#define LRM_POSIT_WARPS 16
#endif
#ifndef LRM_POSIT_CTAS
#define LRM_POSIT_CTAS 3
#endif
constexpr int kWarpsPerCta = LRM_POSIT_WARPS;
#ifndef LRM_POSE_CHUNK
#define LRM_POSE_CHUNK 4
#endif
constexpr int kPoseChunk = LRM_POSE_CHUNK;  // poses a warp takes per visit to the work counter

// ---- the search --------------------------------------------------------------------------------
struct OrientConsts {
    float R[9];        // qtRotate(quat, .) as a matrix (rotateData, several_leg.cu:401-411)
    float radius_in, plus_in, minus_in, radius_out;  // cull cylinders (:505-520)
    float r_near, r_hit;  // world-frame search radii of the two cylinders
    float pad;
    GravExact grav;       // the reference's own rotation sequence, for gravity-plane knife edges
};

struct SearchParams {
    CellGrid map;
    const float* bodies;
    const uint8_t* alive;       // may be nullptr
    size_t nb;
    const OrientConsts* orient; // nq
    const ReachPlan* plans;     // nq * nlegs
    int nq, nlegs;
    int plans_in_smem;
    float r_leg;
    // orientation-independent bounds of the two cull cylinders (rotations keep 3-D distances):
    // a map point nearer than r_collide_all is inside the body cylinder under EVERY orientation, and
    // without a map point within r_near_any the reach cylinder is empty under every orientation
    float r_collide_all, r_near_any;
    // "inside the body cylinder under EVERY orientation", sharper than the ball (leg_math.cuh:
    // AxisCone); cone.ok = 0: ball only; cone.gate = 1: looked for only where the map's cell over the
    // body reaches above it (tilted orientations)
    AxisCone cone;
    uint8_t* standable;
    unsigned long long* next;   // dynamic work counter
    // STATS instantiation only: [0] leg predicates executed (reach_offset on a live map point),
    // [1] cylinder predicates executed, [2] leg predicates of the algorithmic count (every map point
    // inside the reach cylinder, for every leg and orientation, no early exit)
    unsigned long long* stats;
};

__device__ __forceinline__ float3 rotate(const float* R, float x, float y, float z) {
    return make_float3(fmaf(R[0], x, fmaf(R[1], y, R[2] * z)), fmaf(R[3], x, fmaf(R[4], y, R[5] * z)),
                       fmaf(R[6], x, fmaf(R[7], y, R[8] * z)));
}

// Walk over the cells around (bx, by) with three levels of pruning.  The lanes first classify the
// blocks of 4 x 4 cells that overlap the query square with `keep_cell(centre, radius)` — a
// conservative "could hold a witness" test on the block's bounding ball —, then, two surviving
// blocks at a time, their 16 cells each with the same test on the cell's ball, and the warp scans
// the points of the surviving cells 32 at a time.  `visit` returns true to stop.
// Every question asked through this walk is "is there a point such that ...", and the witness is
// usually close to the body: blocks are visited starting with the middle row of the square
// (wrapping around), not from its corner, so that a positive answer comes early.
template <class CellF, class PointF>
__device__ __forceinline__ bool walk_filtered(const CellGrid& g, float bx, float by, float r_xy,
                                              int lane, CellF keep_cell, PointF visit) {
    const int cx0 = max((int)floorf((bx - r_xy - g.x0) * g.inv_cell), 0);
    const int cx1 = min((int)floorf((bx + r_xy - g.x0) * g.inv_cell), g.nx - 1);
    const int cy0 = max((int)floorf((by - r_xy - g.y0) * g.inv_cell), 0);
    const int cy1 = min((int)floorf((by + r_xy - g.y0) * g.inv_cell), g.ny - 1);
    if (cx1 < cx0 || cy1 < cy0) return false;
    const int bx0 = cx0 >> 2, by0 = cy0 >> 2;
    const int w = (cx1 >> 2) - bx0 + 1, h = (cy1 >> 2) - by0 + 1;
    const int nblk = w * h;
    const float cell = 1.0f / g.inv_cell;
    // k / w without an integer division: (k + 0.5) / w is at least 0.5 / w away from an integer,
    // far more than the rounding of the float product (k < 2^20, w < 2^10)
    const float inv_w = 1.0f / (float)w;
    const int first = (h >> 1) * w;
    const int half = lane >> 4, sub = lane & 15;
    for (int base = 0; base < nblk; base += 32) {
        const int kk = base + lane;
        bool keepb = false;
        int gx = 0, gy = 0;  // block coordinates
        if (kk < nblk) {
            const int k = kk + first < nblk ? kk + first : kk + first - nblk;
            const int row = (int)(((float)k + 0.5f) * inv_w);
            gy = by0 + row, gx = bx0 + (k - row * w);
            const float2 ball = g.blk_ball[gy * g.nbx + gx];  // radius < 0: empty block
            if (ball.y >= 0.f)
                keepb = keep_cell(g.x0 + ((float)(4 * gx) + 2.0f) * cell, g.y0 + ((float)(4 * gy) + 2.0f) * cell,
                                  ball.x, ball.y);
        }
        unsigned maskb = __ballot_sync(0xffffffffu, keepb);
        while (maskb) {
            // two blocks at a time: lanes 0-15 take the cells of the first, lanes 16-31 of the second
            const int s0 = __ffs(maskb) - 1;
            maskb &= maskb - 1;
            const int s1 = maskb ? __ffs(maskb) - 1 : -1;
            if (s1 >= 0) maskb &= maskb - 1;
            const int src = half ? (s1 >= 0 ? s1 : s0) : s0;
            const int cx = 4 * __shfl_sync(0xffffffffu, gx, src) + (sub & 3);
            const int cy = 4 * __shfl_sync(0xffffffffu, gy, src) + (sub >> 2);
            bool keep = false;
            int c = 0;
            if ((half == 0 || s1 >= 0) && cx >= cx0 && cx <= cx1 && cy >= cy0 && cy <= cy1) {
                c = cy * g.nx + cx;
                const float2 ball = g.cell_ball[c];  // (z of the centre, radius); radius < 0: empty cell
                if (ball.y >= 0.f)
                    keep = keep_cell(g.x0 + ((float)cx + 0.5f) * cell, g.y0 + ((float)cy + 0.5f) * cell, ball.x, ball.y);
            }
            unsigned mask = __ballot_sync(0xffffffffu, keep);
            while (mask) {
                const int srcc = __ffs(mask) - 1;
                mask &= mask - 1;
                const int cc = __shfl_sync(0xffffffffu, c, srcc);
                const int beg = g.cell_start[cc], end = g.cell_start[cc + 1];
                for (int i = beg; i < end; i += 32) {
                    const int q = i + lane;
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q < end) p = g.pts[q];
                    if (visit(p, q < end)) return true;
                }
            }
        }
    }
    return false;
}

// ---- pre-cull (multi_rot_estimator constructor, several_leg.cu:371-374,413-502) ----------------
// eliminateAlwaysColliding (a map point within 60 mm) and eliminateFarBody (none within 400 mm)
__global__ void body_precull_kernel(CellGrid map, const float* __restrict__ bodies, size_t nb,
                                    uint8_t* __restrict__ alive) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t b = warp; b < nb; b += nwarps) {
        const float bx = bodies[3 * b], by = bodies[3 * b + 1], bz = bodies[3 * b + 2];
        auto within = [&](float r) {
            return walk_filtered(
                map, bx, by, r, lane,
                [&](float x, float y, float z, float rc) { return norm3df(x - bx, y - by, z - bz) < r + rc; },
                [&](float4 t, bool ok) {
                    return __any_sync(0xffffffffu, ok && norm3df(bx - t.x, by - t.y, bz - t.z) < r) != 0;
                });
        };
        const bool keep = !within(60.f) && within(400.f);
        if (lane == 0) alive[b] = keep ? 1 : 0;
    }
}

// eliminateFarTarget: a map point survives if some surviving body lies within 400 mm of it
__global__ void target_precull_kernel(CellGrid body_grid, const float* __restrict__ map, size_t nt,
                                      uint8_t* __restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t t = warp; t < nt; t += nwarps) {
        const float tx = map[3 * t], ty = map[3 * t + 1], tz = map[3 * t + 2];
        const bool found = walk_filtered(
            body_grid, tx, ty, 400.f, lane,
            [&](float x, float y, float z, float rc) { return norm3df(x - tx, y - ty, z - tz) < 400.f + rc; },
            [&](float4 b, bool ok) {
                return __any_sync(0xffffffffu, ok && b.w != 0.f &&
                                                   norm3df(tx - b.x, ty - b.y, tz - b.z) < 400.f) != 0;
            });
        if (lane == 0) keep[t] = found ? 1 : 0;
    }
}

template <bool STATS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, LRM_POSIT_CTAS) positionability_kernel(const SearchParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ReachPlan* plans = P.plans;
    if (P.plans_in_smem) {
        // all (orientation, leg) plans live in shared memory for the whole kernel
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        const float4* src = reinterpret_cast<const float4*>(P.plans);
        const int n16 = P.nq * P.nlegs * (int)(sizeof(ReachPlan) / 16);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        plans = reinterpret_cast<const ReachPlan*>(smem_raw);
    }
    const int lane = threadIdx.x & 31;
    unsigned long long n_leg = 0, n_cyl = 0, n_alg = 0;  // STATS: identical in every lane of the warp
    // Work is handed out by one atomic counter, kPoseChunk consecutive poses at a time: a counter
    // bumped once per pose serialises 16.8 M same-address atomics (~24 ns each on a B200: the whole
    // 0.41 s of configs[2]); consecutive poses are neighbours in space and cost about the same.
    unsigned long long b = 0, b_end = 0;
    while (true) {
        if (b == b_end) {
            if (lane == 0) b = atomicAdd(P.next, (unsigned long long)kPoseChunk);
            b = __shfl_sync(0xffffffffu, b, 0);
            b_end = b + kPoseChunk < P.nb ? b + kPoseChunk : P.nb;
            if (b >= P.nb) {
                if (STATS && lane == 0) {
                    atomicAdd(&P.stats[0], n_leg), atomicAdd(&P.stats[1], n_cyl), atomicAdd(&P.stats[2], n_alg);
                }
                return;
            }
        }

        uint8_t result = 0;
        if (P.alive == nullptr || P.alive[b]) {
            const float bx = P.bodies[3 * b], by = P.bodies[3 * b + 1], bz = P.bodies[3 * b + 2];
            // One orientation-independent scan settles most poses of a dense pose lattice (far above
            // the map, or inside it) before the per-orientation scans: exact, because both bounds
            // are conservative for every rotation.
            bool collide_all = walk_filtered(
                P.map, bx, by, P.r_collide_all, lane,
                [&](float x, float y, float z, float rc) {
                    return norm3df(x - bx, y - by, z - bz) < P.r_collide_all + rc;
                },
                [&](float4 t, bool ok) {
                    const float d = norm3df(t.x - bx, t.y - by, t.z - bz);
                    return __any_sync(0xffffffffu, ok && t.w != 0.f && d < P.r_collide_all) != 0;
                });
            // The ball misses a body that sits UNDER the terrain by more than its radius: the map is a
            // surface, and the points above such a body are 110 .. 250 mm away, yet inside the cylinder
            // whatever the tilt.  With tilted orientations that is all the region adds, so it is looked
            // for only where the map's cell over the body reaches above it (a pose above the ground has
            // nothing to find there, and on a dense map the wider window costs a walk: configs[3] 325
            // -> 292 ms).  With level orientations only (all cylinder axes equal: configs[4]) the region IS
            // the cylinder, and one walk settles what every orientation would find: always taken
            // (42 - 46 ms against 50).
            if (!collide_all && P.cone.ok) {
                const int cx = (int)floorf((bx - P.map.x0) * P.map.inv_cell), cy = (int)floorf((by - P.map.y0) * P.map.inv_cell);
                bool above = !P.cone.gate;
                if (P.cone.gate && cx >= 0 && cy >= 0 && cx < P.map.nx && cy < P.map.ny) {
                    const int c = cy * P.map.nx + cx;
                    above = P.map.cell_start[c + 1] == P.map.cell_start[c] || P.map.cell_z[c].y > bz;  // empty cell: unknown
                }
                if (above)
                    collide_all = walk_filtered(
                        P.map, bx, by, P.cone.hi, lane,
                        [&](float x, float y, float z, float rc) {
                            // a superset of the upper part of the region: the ball of the cylinder's
                            // height, above the body's plane
                            const float dx = x - bx, dy = y - by, dz = z - bz;
                            return norm3df(dx, dy, dz) < P.cone.hi + rc &&
                                   fmaf(P.cone.ax, dx, fmaf(P.cone.ay, dy, P.cone.az * dz)) + rc > (P.cone.gate ? 0.f : P.cone.lo);
                        },
                        [&](float4 t, bool ok) {
                            return __any_sync(0xffffffffu, ok && t.w != 0.f && cone_collides_always(P.cone, t.x - bx, t.y - by, t.z - bz)) != 0;
                        });
            }
            const bool near_any = !collide_all && walk_filtered(
                P.map, bx, by, P.r_near_any, lane,
                [&](float x, float y, float z, float rc) {
                    return norm3df(x - bx, y - by, z - bz) < P.r_near_any + rc;
                },
                [&](float4 t, bool ok) {
                    const float d = norm3df(t.x - bx, t.y - by, t.z - bz);
                    return __any_sync(0xffffffffu, ok && t.w != 0.f && d < P.r_near_any) != 0;
                });
            const int nq = (collide_all || !near_any) ? 0 : P.nq;
            // An orientation passes when no map point is inside the body cylinder, one is inside the
            // reach cylinder, and every leg reaches one: an AND, so the order of the tests is free.
            // A pose that cannot stand usually fails the same test under the next orientation:
            // the legs are tried starting with the one that failed last (a failing walk has no
            // early exit), and when a leg failed last that leg goes before the two cylinder walks.
            int first_leg = 0;
            bool leg_failed_last = false;
            for (int o = 0; o < nq && result == 0; o++) {
                const OrientConsts& O = P.orient[o];
                const float3 B = rotate(O.R, bx, by, bz);
                bool pass = true;
                for (int step = 0; step < 2 + P.nlegs && pass; step++) {
                    // leg first: [failed leg, body cylinder, reach cylinder, other legs]; else
                    // [body cylinder, reach cylinder, legs from the one that failed last]
                    const int what = leg_failed_last ? (step == 0 ? 2 : (step <= 2 ? step - 1 : step)) : step;
                    if (what == 0) {
                        // eliminateFarAndColliding (several_leg.cu:504-559): no map point inside the
                        // body cylinder (r = dim.body, z in (-110, 250)) ...
                        pass = !walk_filtered(
                            P.map, bx, by, O.r_hit, lane,
                            [&](float x, float y, float z, float rc) {
                                const float3 T = rotate(O.R, x, y, z);
                                const float dz = T.z - B.z;
                                const float dx = T.x - B.x, dy = T.y - B.y, rr = O.radius_out + rc;
                                return fmaf(dx, dx, dy * dy) < rr * rr && dz < 250.f + rc && dz > -110.f - rc;
                            },
                            [&](float4 t, bool ok) {
                                const float3 T = rotate(O.R, t.x, t.y, t.z);
                                const float dz = T.z - B.z;
                                const float dx = T.x - B.x, dy = T.y - B.y;
                                const bool in_body = fmaf(dx, dx, dy * dy) < O.radius_out * O.radius_out &&
                                                     dz < 250.f && dz > -110.f;
                                if (STATS) n_cyl += __popc(__ballot_sync(0xffffffffu, ok && t.w != 0.f));
                                return __any_sync(0xffffffffu, ok && t.w != 0.f && in_body) != 0;
                            });
                        if (!pass) leg_failed_last = false;
                    } else if (what == 1) {
                        // ... and at least one inside the reach cylinder
                        pass = walk_filtered(
                            P.map, bx, by, O.r_near, lane,
                            [&](float x, float y, float z, float rc) {
                                const float3 T = rotate(O.R, x, y, z);
                                const float dz = T.z - B.z;
                                const float dx = T.x - B.x, dy = T.y - B.y, rr = O.radius_in + rc;
                                return fmaf(dx, dx, dy * dy) < rr * rr && dz < O.plus_in + rc && dz > O.minus_in - rc;
                            },
                            [&](float4 t, bool ok) {
                                const float3 T = rotate(O.R, t.x, t.y, t.z);
                                const float dz = T.z - B.z;
                                const float dx = T.x - B.x, dy = T.y - B.y;
                                const bool in_reach = fmaf(dx, dx, dy * dy) < O.radius_in * O.radius_in &&
                                                      dz < O.plus_in && dz > O.minus_in;
                                if (STATS) n_cyl += __popc(__ballot_sync(0xffffffffu, ok && t.w != 0.f));
                                return __any_sync(0xffffffffu, ok && t.w != 0.f && in_reach) != 0;
                            });
                        if (!pass) leg_failed_last = false;
                    } else {
                        // eliminateUnreachable (:633-706): every leg needs one reachable map point
                        const int ll = what - 2;  // 0: the leg that failed last
                        const int l = first_leg + ll < P.nlegs ? first_leg + ll : first_leg + ll - P.nlegs;
                        const ReachPlan& L = plans[o * P.nlegs + l];
                        pass = walk_filtered(
                            P.map, bx, by, P.r_leg, lane,
                            [&](float x, float y, float z, float rc) {
                                const float3 T = rotate(O.R, x, y, z);
                                return reach_ball_possible(L, T.x - B.x, T.y - B.y, T.z - B.z, rc);
                            },
                            [&](float4 t, bool ok) {
                                // reachable_rotate_leg (several_leg.cu:48-67): offset in the
                                // orientation frame, gravity-side test, leg frame, reachability_circles
                                const float3 T = rotate(O.R, t.x, t.y, t.z);
                                const GravCtx gc{&O.grav, bx, by, bz, t.x, t.y, t.z};
                                const bool r = ok && t.w != 0.f && reach_offset(L, T.x - B.x, T.y - B.y, T.z - B.z, &gc);
                                if (STATS) n_leg += __popc(__ballot_sync(0xffffffffu, ok && t.w != 0.f));
                                return __any_sync(0xffffffffu, r) != 0;
                            });
                        if (!pass) first_leg = l, leg_failed_last = true;
                    }
                }
                if (pass) result = (uint8_t)(o + 1);
            }
            if (STATS) {
                // the algorithmic count: what a search without pruning or early exit evaluates — for
                // every orientation, every map point inside the reach cylinder, once per leg
                for (int o = 0; o < P.nq; o++) {
                    const OrientConsts& O = P.orient[o];
                    const float3 B = rotate(O.R, bx, by, bz);
                    unsigned long long inside = 0;
                    walk_filtered(
                        P.map, bx, by, O.r_near, lane,
                        [&](float x, float y, float z, float rc) {
                            const float3 T = rotate(O.R, x, y, z);
                            const float dz = T.z - B.z;
                            const float dx = T.x - B.x, dy = T.y - B.y, rr = O.radius_in + rc;
                            return fmaf(dx, dx, dy * dy) < rr * rr && dz < O.plus_in + rc && dz > O.minus_in - rc;
                        },
                        [&](float4 t, bool ok) {
                            const float3 T = rotate(O.R, t.x, t.y, t.z);
                            const float dz = T.z - B.z;
                            const float dx = T.x - B.x, dy = T.y - B.y;
                            const bool in_reach = fmaf(dx, dx, dy * dy) < O.radius_in * O.radius_in &&
                                                  dz < O.plus_in && dz > O.minus_in;
                            inside += __popc(__ballot_sync(0xffffffffu, ok && t.w != 0.f && in_reach));
                            return false;
                        });
                    n_alg += inside * (unsigned long long)P.nlegs;
                }
            }
        }
        if (lane == 0) P.standable[b] = result;
        b++;
    }
}

// ---- host orchestration ------------------------------------------------------------------------
}  // namespace

cudaError_t run_positionability(const PositParams& p, cudaStream_t stream, float* kernel_ms) {
    if (p.nb == 0) return cudaSuccess;
    if (p.nt > 0x7fffffffull || p.nb > 0x7fffffffull * 64) return cudaErrorInvalidValue;
    DevBuf mem;
    // ~48 map points per cell: fine enough for the cell-level pruning to bite, coarse enough that
    // a query touches a few hundred cells
    const float map_cell = 0.f; // search grid: sized from the map density inside build_grid

    // per-orientation / per-leg constants
    std::vector<OrientConsts> orient(p.nq);
    std::vector<ReachPlan> plans((size_t)p.nq * p.nlegs);
    float r_leg = 0.f;
    const float pi = 3.14159265358979323846264338327950288419716939937510582097f;
    for (int o = 0; o < p.nq; o++) {
        const float* q = p.quats + 4 * o;
        const float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
        if (!(std::fabs(n2 - 1.f) < 1e-3f)) return cudaErrorInvalidValue;  // must be a rotation
        const float ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
        float c0[3], c1[3], c2[3];
        quat_rotate(q, ex, c0), quat_rotate(q, ey, c1), quat_rotate(q, ez, c2);
        OrientConsts& O = orient[o];
        for (int r = 0; r < 3; r++) O.R[3 * r] = c0[r], O.R[3 * r + 1] = c1[r], O.R[3 * r + 2] = c2[r];
        for (int l = 0; l < p.nlegs; l++) {
            LegPlan full;
            build_leg_plan_rotated_limits(p.legs[l], q, &full);
            float az_s, az_c;
            sincosf(-p.legs[l].body_angle, &az_s, &az_c);  // rotateInPlace, several_leg.cu:26-32
            make_reach_plan(full, p.legs[l].min_angle_coxa, p.legs[l].max_angle_coxa, az_c, az_s,
                            &plans[(size_t)o * p.nlegs + l]);
            const lrm_leg_t& d = p.legs[l];
            r_leg = std::fmax(r_leg, std::fabs(d.body) + std::fabs(d.coxa_length) +
                                         std::fabs(d.femur_length) + std::fabs(d.tibia_length) + 1.f);
        }
        // cull cylinders from leg 0 after the limit rotation (several_leg.cu:505-520)
        lrm_leg_t d = p.legs[0];
        const float pitch = quat_pitch_for_leg(q, d.body_angle);
        d.tibia_absolute_pos -= pitch, d.tibia_absolute_neg -= pitch;
        const float s_p = std::sin(d.coxa_pitch), c_p = std::cos(d.coxa_pitch);
        O.radius_in = d.body + c_p * d.coxa_length + d.femur_length + d.tibia_length;
        const float plus_abs = d.tibia_length * std::sin(d.tibia_absolute_pos) +
                               d.femur_length * std::sin(std::min(pi / 2, d.max_angle_femur));
        O.plus_in = s_p * d.coxa_length + plus_abs;
        O.minus_in = s_p * d.coxa_length - d.femur_length - d.tibia_length;
        O.radius_out = d.body;
        // a rotation keeps 3-D distances: every point of a cylinder lies within this world radius
        const float zin = std::fmax(std::fabs(O.plus_in), std::fabs(O.minus_in));
        O.r_near = std::sqrt(O.radius_in * O.radius_in + zin * zin) + 1.f;
        O.r_hit = std::sqrt(O.radius_out * O.radius_out + 250.f * 250.f) + 1.f;
        O.pad = 0.f;
        make_grav_exact(q, &O.grav);
    }

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (kernel_ms) {
        POSIT_CHECK(cudaEventCreate(&ev0));
        POSIT_CHECK(cudaEventCreate(&ev1));
        POSIT_CHECK(cudaEventRecord(ev0, stream));
    }
    auto finish = [&](cudaError_t e) {
        if (kernel_ms && e == cudaSuccess) {
            e = cudaEventRecord(ev1, stream);
            if (e == cudaSuccess) e = cudaEventSynchronize(ev1);
            if (e == cudaSuccess) e = cudaEventElapsedTime(kernel_ms, ev0, ev1);
        }
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        return e;
    };

    uint8_t* alive = nullptr;
    uint8_t* keep = nullptr;
    CellGrid map_grid;
    if (p.pre_cull) {
        CellGrid raw;
        cudaError_t e = build_grid(mem, p.map, p.nt, nullptr, map_cell, stream, &raw);
        if (e != cudaSuccess) return finish(e);
        if ((e = mem.alloc(&alive, p.nb)) != cudaSuccess) return finish(e);
        if ((e = mem.alloc(&keep, p.nt)) != cudaSuccess) return finish(e);
        body_precull_kernel<<<148 * 8, 256, 0, stream>>>(raw, p.bodies, p.nb, alive);
        CellGrid body_grid;
        if ((e = build_grid(mem, p.bodies, p.nb, alive, map_cell, stream, &body_grid)) != cudaSuccess)
            return finish(e);
        target_precull_kernel<<<148 * 8, 256, 0, stream>>>(body_grid, p.map, p.nt, keep);
        if ((e = build_grid(mem, p.map, p.nt, keep, map_cell, stream, &map_grid)) != cudaSuccess)
            return finish(e);
    } else {
        cudaError_t e = build_grid(mem, p.map, p.nt, nullptr, map_cell, stream, &map_grid);
        if (e != cudaSuccess) return finish(e);
    }

    OrientConsts* d_orient;
    ReachPlan* d_plans;
    unsigned long long* d_next;
    cudaError_t e;
    if ((e = mem.alloc(&d_orient, orient.size())) != cudaSuccess) return finish(e);
    if ((e = mem.alloc(&d_plans, plans.size())) != cudaSuccess) return finish(e);
    if ((e = mem.alloc(&d_next, 1)) != cudaSuccess) return finish(e);
    cudaMemcpyAsync(d_orient, orient.data(), orient.size() * sizeof(OrientConsts), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_plans, plans.data(), plans.size() * sizeof(ReachPlan), cudaMemcpyHostToDevice, stream);
    cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), stream);

    SearchParams S;
    S.map = map_grid, S.bodies = p.bodies, S.alive = alive, S.nb = p.nb;
    S.orient = d_orient, S.plans = d_plans, S.nq = p.nq, S.nlegs = p.nlegs;
    S.r_leg = r_leg;
    {
        // body cylinder (r = radius_out, z in (-110, 250)): the ball of radius min(r, 110) around the
        // body centre lies inside it; reach cylinder: contained in the ball of radius r_near.
        float r_in = 1.0e30f, r_far = 0.f;
        for (const OrientConsts& O : orient) {
            r_in = std::fmin(r_in, std::fmin(O.radius_out, 110.f));
            r_far = std::fmax(r_far, O.r_near);
        }
        S.r_collide_all = std::fmax(0.f, r_in - 1.f);  // 1 mm of slack for the rounding of the rotations
        S.r_near_any = r_far + 1.f;
        // the cone of cylinder axes (third row of R: height = R[6..8] . d)
        std::vector<float> axes;
        float r_body = 1.0e30f;
        for (const OrientConsts& O : orient) {
            axes.push_back(O.R[6]), axes.push_back(O.R[7]), axes.push_back(O.R[8]);
            r_body = std::fmin(r_body, O.radius_out);
        }
        make_axis_cone(axes.data(), (int)orient.size(), -110.f, 250.f, r_body, &S.cone);
#ifdef LRM_NO_CONE
        S.cone.ok = 0;
#endif
    }
    S.standable = p.standable, S.next = d_next;
    S.stats = nullptr;
    if (p.stats) {
        if ((e = mem.alloc(&S.stats, 3)) != cudaSuccess) return finish(e);
        cudaMemsetAsync(S.stats, 0, 3 * sizeof(unsigned long long), stream);
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // plans in shared memory when they fit twice per SM (two CTAs of 16 warps), else from L1/L2
    const size_t plan_bytes = plans.size() * sizeof(ReachPlan);
    S.plans_in_smem = plan_bytes <= 100 * 1024 ? 1 : 0;
    const size_t smem = S.plans_in_smem ? plan_bytes : 0;
    auto kernel = p.stats ? positionability_kernel<true> : positionability_kernel<false>;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return finish(e);
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kWarpsPerCta * 32, smem);
    if (occ < 1) occ = 1;
    kernel<<<sms * occ, kWarpsPerCta * 32, smem, stream>>>(S);
    e = cudaGetLastError();
    if (e != cudaSuccess) return finish(e);
    e = finish(cudaSuccess);
    if (e != cudaSuccess) return e;
    if (p.stats) {
        unsigned long long h[3] = {0, 0, 0};
        e = cudaMemcpyAsync(h, S.stats, sizeof h, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        for (int i = 0; i < 3; i++) p.stats[i] = (double)h[i];
        return e;
    }
    // scratch (grid, plans) is freed on return: make sure the kernels are done with it
    return cudaStreamSynchronize(stream);
}

}  // namespace lrm
