// leg_math.cuh — per-point device math of the one-leg reachability / distance path.
//
// Same decisions and the same nearest-boundary construction as the reference
// (reachability_circles one_leg.cu:280-319, distance_circles :321-341, finish_finding_closest
// :215-278, eval_plane_circles :167-208, multi_circle_clamp :91-145, find_region
// circles.cu.h:48-78), re-derived so that nothing leg-constant is evaluated per point:
//   * no atan2f / sincosf: every angle comparison is a cross-product sign test against a constant
//     direction (AngleTest), and the coxa rotation uses the normalised (x, y) itself;
//   * the circles of a sector come from a host-built 4-sector table staged in shared memory
//     (two float4 per circle) instead of being rebuilt from 8+ sin/cos per point into local memory;
//   * "does the projection on circle j satisfy the other circles" (12 circle tests per plane
//     evaluation in the reference) is one dot product against the precomputed valid arc of
//     circle j; corner points are constants;
//   * the direct and the pi-flipped coxa solution share their yaw tests, and a solution whose
//     plane evaluation provably duplicates the other one's (yaw beyond limit +- pi/2) is skipped.
// Everything lives in registers; the only memory traffic is the point itself.
#pragma once
#include <cuda_runtime.h>

#include "leg_plan.h"

// The math below is plain FP32 and compiles for the host as well: tests/emu builds it into a
// test-only emulator so the CPU test suite can check this file against the oracle without a GPU.
// The product library only ever instantiates the __device__ side.
#define LRM_HD __host__ __device__ __forceinline__

namespace lrm {

LRM_HD int f2i(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i;
    __builtin_memcpy(&i, &f, 4);
    return i;
#endif
}
// 1/sqrt(x), ~2 ulp, one MUFU.RSQ (denormal inputs flush to zero -> +inf, handled by callers)
LRM_HD float fast_rsqrt(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}

// 1/x, ~1 ulp, one MUFU.RCP
LRM_HD float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

constexpr float kMarginF = 0.001f;  // CIRCLE_MARGIN, settings.h:9

// Sector table staged in shared memory by every CTA: a copy of LegPlan::sector.
// row[sector][2*j]   = (cx, cy, r, sgn)      of slot j+1
// row[sector][2*j+1] = (ax, ay, ah, inner_j) valid arc of slot j+1; 4th = j-th arc number of slot 0
// row[sector][6]     = second valid arc of slot 0 (the inner circle)
// Row stride 7 float4 = 28 words: lanes in different sectors hit different banks.
struct SectorTable {
    float4 row[4][7];
};

LRM_HD void fill_sector_table(const LegPlan& L, SectorTable* tab, int tid, int nthreads) {
    for (int i = tid; i < 28; i += nthreads) {
        const int s = i / 7, k = i % 7;
        const float* src = k < 6 ? &L.sector[s].slot[k >> 1][(k & 1) * 4] : L.sector[s].inner_b;
        tab->row[s][k] = make_float4(src[0], src[1], src[2], src[3]);
    }
}

// 1.0f when atan2f(Y, .) lies in [0, pi] (sign bit of Y clear), else 0.0f
LRM_HD float up_flag(float Y) { return f2i(Y) >= 0 ? 1.f : 0.f; }

// atan2f(Y, X) > theta, see AngleTest: 3 FFMA + 1 FSETP, no predicate logic
LRM_HD bool angle_gt(const AngleTest& t, float X, float Y, float upf) {
    const float cr = fmaf(t.c, Y, fmaf(t.ns, X, t.bias));
    return cr > fmaf(upf, -kAngleBig, t.thr_dn);
}
LRM_HD bool angle_gt(const AngleTest& t, float X, float Y) { return angle_gt(t, X, Y, up_flag(Y)); }

LRM_HD int find_sector(const LegPlan& L, float X, float Y) {
    const float upf = up_flag(Y);
    const bool upper = angle_gt(L.middle, X, Y, upf);
    const bool more0 = angle_gt(L.sat[0], X, Y, upf), more1 = angle_gt(L.sat[1], X, Y, upf);
    const bool more = upper ? more1 : more0;
    const bool ext = upper != more;  // circles.cu.h:73-74
    return (upper ? 2 : 0) | (ext ? 1 : 0);
}

// validity of a point against one circle on squared distances (one_leg.cu:31-41):
// attractive: |v| < r + eps, repulsive: |v| > r - eps
LRM_HD bool circle_ok(float cx, float cy, float r, float sgn, float x, float y) {
    const float vx = x - cx, vy = y - cy;
    const float t = fmaf(sgn, kMarginF, r);
    return sgn * fmaf(vx, vx, vy * vy) < sgn * t * t;
}

// eval_plane_circles<REACH_USECASE>, (X, Y) already relative to the femur joint.
LRM_HD bool plane_reach(const LegPlan& L, const SectorTable& tab, float X, float Y) {
    const int s = find_sector(L, X, Y);
    bool ok = L.inner.sgn * fmaf(X, X, Y * Y) < L.inner.thr_s;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 c = tab.row[s][2 * j];
        ok = ok & circle_ok(c.x, c.y, c.z, c.w, X, Y);
    }
    return ok;
}

struct PlaneResult {
    bool valid;    // the query point satisfies all 4 circles
    float dx, dy;  // P - nearest valid boundary candidate
};

// eval_plane_circles<DIST_USECASE> = insert_circles + insert_intersecv2 + multi_circle_clamp.
// GENERIC = false: projections are validated against the precomputed arcs.
// GENERIC = true : explicit cross-validation against the other three circles (any leg).
template <bool GENERIC>
LRM_HD PlaneResult plane_clamp(const LegPlan& L, const SectorTable& tab, float X, float Y) {
    const int s = find_sector(L, X, Y);
    float cx[4], cy[4], r[4], sg[4], ax[4], ay[4], ah[4];
    cx[0] = 0.f, cy[0] = 0.f, r[0] = L.inner.r, sg[0] = L.inner.sgn;
    const float4 inner_b = tab.row[s][6];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 c = tab.row[s][2 * j];
        const float4 a = tab.row[s][2 * j + 1];
        cx[j + 1] = c.x, cy[j + 1] = c.y, r[j + 1] = c.z, sg[j + 1] = c.w;
        ax[j + 1] = a.x, ay[j + 1] = a.y, ah[j + 1] = a.z;
        if (j == 0) ax[0] = a.w;
        if (j == 1) ay[0] = a.w;
        if (j == 2) ah[0] = a.w;
    }

    // project P on each circle (force_clamp_on_circle, one_leg.cu:42-63); a projection only counts
    // if it satisfies the other circles (:122-123); the first strictly closer one wins (:133-140)
    float px[4], py[4], d[4];
    bool cand[4];
    bool valid = true;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float vx = X - cx[j], vy = Y - cy[j];
        const float m2 = fmaf(vx, vx, vy * vy);
        float rinv = fast_rsqrt(m2);
        float m = m2 * rinv;
        float len = m;
        if (__builtin_expect(!(m >= kMarginF), 0)) {  // P on the centre (m2 == 0: m = NaN)
            m = m2 > 0.f ? m : 0.f;
            vx = 1.f, vy = 0.f, rinv = 1.f, len = 1.f;
        }
        d[j] = r[j] - m;
        valid = valid & (sg[j] * d[j] > -kMarginF);  // (d >= 0) == attractive, or |d| < margin
        const float k = r[j] * rinv;
        px[j] = fmaf(vx, k, cx[j]);
        py[j] = fmaf(vy, k, cy[j]);
        // direction of the projection inside the valid arc of circle j
        cand[j] = fmaf(vx, ax[j], vy * ay[j]) >= ah[j] * len;
        if (j == 0) cand[0] = cand[0] | (fmaf(vx, inner_b.x, vy * inner_b.y) >= inner_b.z * len);
    }
    if (GENERIC) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k != j) ok = ok & circle_ok(cx[k], cy[k], r[k], sg[k], px[j], py[j]);
            cand[j] = ok;
        }
    }
    float best_abs = 999999999999999.9f, bx = 0.f, by = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float a = fabsf(d[j]);
        const bool take = cand[j] & (best_abs > a);
        best_abs = take ? a : best_abs;
        bx = take ? px[j] : bx;
        by = take ? py[j] : by;
    }
    // corner points compete only when P itself is outside (one_leg.cu:109-118)
    if (!valid) {
        float best2 = best_abs * best_abs;
#pragma unroll
        for (int i = 0; i < kMaxCorners; i++) {
            if (i < L.n_corners) {
                const float wx = X - L.corner_x[i], wy = Y - L.corner_y[i];
                const float w2 = fmaf(wx, wx, wy * wy);
                const bool take = best2 > w2;
                best2 = take ? w2 : best2;
                bx = take ? L.corner_x[i] : bx;
                by = take ? L.corner_y[i] : by;
            }
        }
    }
    PlaneResult out;
    out.valid = valid;
    out.dx = X - bx;
    out.dy = Y - by;
    return out;
}

// ---- compact plan for the positionability search ------------------------------------------------
// Only what reachable_rotate_leg (several_leg.cu:48-67) needs, 16-byte aligned so that a few
// hundred (orientation, leg) plans fit in one CTA's shared memory.
struct alignas(16) ReachPlan {
    float4 circ[4][3];  // [sector][slot-1] = (cx, cy, r, sgn)
    float M[9], t[3];   // orientation-frame foothold offset -> coxa frame
    float grav[3];      // gravity-side half-space (several_leg.cu:58-62)
    float coxa_length;
    AngleTest over, under, middle, sat[2];
    float inner_sgn, inner_thr_s;
    float r_min, r_max;      // inner / outer circle radii   (cell-level pruning)
    float yaw_min, yaw_max;  // coxa yaw limits, radians     (cell-level pruning)
    float az_cos, az_sin;    // sincosf(-body_angle) as the reference evaluates it (knife-edge recheck)
    // unit directions of the two yaw limits; wedge != 0 when the limits span less than pi, i.e.
    // the allowed yaws are the intersection of two half-planes through the coxa axis (cell pruning)
    float cmin, smin, cmax, smax, wedge;
};

LRM_HD void make_reach_plan(const LegPlan& L, float yaw_min, float yaw_max, float az_cos, float az_sin,
                            ReachPlan* R) {
    for (int s = 0; s < 4; s++)
        for (int j = 0; j < 3; j++) {
            const float* o = L.sector[s].slot[j];
            R->circ[s][j] = make_float4(o[0], o[1], o[2], o[3]);
        }
    for (int i = 0; i < 9; i++) R->M[i] = L.M[i];
    for (int i = 0; i < 3; i++) R->t[i] = L.t[i], R->grav[i] = L.grav[i];
    R->coxa_length = L.coxa_length;
    R->over = L.over, R->under = L.under, R->middle = L.middle, R->sat[0] = L.sat[0], R->sat[1] = L.sat[1];
    R->inner_sgn = L.inner.sgn, R->inner_thr_s = L.inner.thr_s;
    R->r_min = L.inner.r, R->r_max = L.outer.r;
    R->yaw_min = yaw_min, R->yaw_max = yaw_max;
    R->az_cos = az_cos, R->az_sin = az_sin;
    R->cmin = cosf(yaw_min), R->smin = sinf(yaw_min), R->cmax = cosf(yaw_max), R->smax = sinf(yaw_max);
    R->wedge = (yaw_max >= yaw_min && yaw_max - yaw_min < 3.0f) ? 1.f : 0.f;
}

// ---- the gravity-side test on its knife edge -----------------------------------------------------
// reachable_rotate_leg rejects a foothold when x < 0 for
//     x = Rz(-body_angle) * qtRotate(qtInvert(q), qtRotate(q, t) - qtRotate(q, b))
// (several_leg.cu:48-62, rotateData :401-411), i.e. for the WORLD offset t - b up to the rounding
// of three rotations.  When the pose lattice coincides with the map lattice whole columns of
// footholds have t.x - b.x = 0 exactly and the reference's decision is that rounding.  The search
// evaluates the test with one fused dot product; within kGravBand of zero it re-evaluates it
// with the reference's own sequence of individually rounded operations, so that even those
// decisions are the reference's.
struct GravExact {
    float rq[9];   // coefficient sums of qtRotate(q, .), unified_math_cuda.cu.h:13-27, in its order
    float rqi[9];  // the same for qtInvert(q)
};
LRM_HD float nf_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);  // never contracted into an FMA
#else
    return a * b;            // host emulation is built with -ffp-contract=off
#endif
}
LRM_HD float nf_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
// 2.0f * ((c0*v.x + c1*v.y) + c2*v.z) + w
LRM_HD float ref_rot_row(const float* c, float x, float y, float z, float w) {
    return nf_add(nf_mul(2.0f, nf_add(nf_add(nf_mul(c[0], x), nf_mul(c[1], y)), nf_mul(c[2], z))), w);
}
LRM_HD void make_grav_exact(const float q[4], GravExact* G) {
    auto fill = [](float x, float y, float z, float w, float* c) {
        const float t2 = x * y, t3 = x * z, t4 = x * w, t5 = -y * y, t6 = y * z, t7 = y * w, t8 = -z * z,
                    t9 = z * w, t10 = -w * w;
        c[0] = t8 + t10, c[1] = t6 - t4, c[2] = t3 + t7;
        c[3] = t4 + t6, c[4] = t5 + t10, c[5] = t9 - t2;
        c[6] = t7 - t3, c[7] = t2 + t9, c[8] = t5 + t8;
    };
    fill(q[0], q[1], q[2], q[3], G->rq);
    const float n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];  // qtInvert, :29-34
    fill(q[0] / n, -q[1] / n, -q[2] / n, -q[3] / n, G->rqi);
}
struct GravCtx {
    const GravExact* G;
    float bx, by, bz;  // body position, world frame
    float tx, ty, tz;  // foothold, world frame
};
constexpr float kGravBand = 0.02f, kGravBandRel = 4.0e-6f;
LRM_HD bool grav_rejects_exact(const ReachPlan& L, const GravCtx& c) {
    const float* q = c.G->rq;
    const float Bx = ref_rot_row(q, c.bx, c.by, c.bz, c.bx), By = ref_rot_row(q + 3, c.bx, c.by, c.bz, c.by),
                Bz = ref_rot_row(q + 6, c.bx, c.by, c.bz, c.bz);
    const float Tx = ref_rot_row(q, c.tx, c.ty, c.tz, c.tx), Ty = ref_rot_row(q + 3, c.tx, c.ty, c.tz, c.ty),
                Tz = ref_rot_row(q + 6, c.tx, c.ty, c.tz, c.tz);
    const float vx = nf_add(Tx, -Bx), vy = nf_add(Ty, -By), vz = nf_add(Tz, -Bz);
    const float* qi = c.G->rqi;
    const float gx = ref_rot_row(qi, vx, vy, vz, vx), gy = ref_rot_row(qi + 3, vx, vy, vz, vy);
    // rotateInPlace(gravity_down, -body_angle): x * cos - y * sin
    return nf_add(nf_mul(gx, L.az_cos), -nf_mul(gy, L.az_sin)) < 0.f;
}

// reachable_rotate_leg for a foothold offset (vx, vy, vz) in the orientation frame:
// gravity-side test, leg frame, reachability_circles.  Same arithmetic as
// to_coxa_frame + reach_coxa_frame on the full plan.
LRM_HD bool reach_offset(const ReachPlan& L, float vx, float vy, float vz, const GravCtx* ctx = nullptr) {
    const float g = fmaf(L.grav[0], vx, fmaf(L.grav[1], vy, L.grav[2] * vz));
    if (ctx != nullptr &&
        fabsf(g) < fmaf(kGravBandRel, fabsf(ctx->bx) + fabsf(ctx->by) + fabsf(ctx->bz) + fabsf(ctx->tx) +
                                          fabsf(ctx->ty) + fabsf(ctx->tz), kGravBand)) {
        if (grav_rejects_exact(L, *ctx)) return false;
    } else if (g < 0.f) {
        return false;
    }
    const float px = fmaf(L.M[0], vx, fmaf(L.M[1], vy, fmaf(L.M[2], vz, L.t[0])));
    const float py = fmaf(L.M[3], vx, fmaf(L.M[4], vy, fmaf(L.M[5], vz, L.t[1])));
    const float pz = fmaf(L.M[6], vx, fmaf(L.M[7], vy, fmaf(L.M[8], vz, L.t[2])));
    const bool flip = f2i(px) < 0;
    const float xf = flip ? -px : px, yf = flip ? -py : py;
    if (angle_gt(L.over, xf, yf) | angle_gt(L.under, xf, -yf)) return false;
    const float rho2 = fmaf(px, px, py * py);
    const float rho = rho2 > 0.f ? rho2 * fast_rsqrt(rho2) : 0.f;
    const float X = (flip ? -rho : rho) - L.coxa_length, Y = pz;
    const float upf = up_flag(Y);
    const bool upper = angle_gt(L.middle, X, Y, upf);
    const bool more = upper ? angle_gt(L.sat[1], X, Y, upf) : angle_gt(L.sat[0], X, Y, upf);
    const int s = (upper ? 2 : 0) | ((upper != more) ? 1 : 0);
    bool ok = L.inner_sgn * fmaf(X, X, Y * Y) < L.inner_thr_s;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 c = L.circ[s][j];
        ok = ok & circle_ok(c.x, c.y, c.z, c.w, X, Y);
    }
    return ok;
}

// Can ANY point within `rc` of the offset v be reachable?  Conservative (never rejects a ball that
// holds a reachable point): gravity half-space, the annulus between the inner and outer circle in
// the femur plane, and the yaw wedge, each widened by rc; both coxa solutions are considered.
LRM_HD bool reach_ball_possible(const ReachPlan& L, float vx, float vy, float vz, float rc) {
    const float g = fmaf(L.grav[0], vx, fmaf(L.grav[1], vy, L.grav[2] * vz));
    if (g < -rc) return false;
    const float px = fmaf(L.M[0], vx, fmaf(L.M[1], vy, fmaf(L.M[2], vz, L.t[0])));
    const float py = fmaf(L.M[3], vx, fmaf(L.M[4], vy, fmaf(L.M[5], vz, L.t[1])));
    const float pz = fmaf(L.M[6], vx, fmaf(L.M[7], vy, fmaf(L.M[8], vz, L.t[2])));
    const float rho2 = fmaf(px, px, py * py);
    const float rho = sqrtf(rho2);
    const float lo = fmaxf(L.r_min - rc - 0.01f, 0.f), hi = L.r_max + rc + 0.01f;
    const float xa = rho - L.coxa_length, xb = -rho - L.coxa_length;
    const float da2 = fmaf(xa, xa, pz * pz), db2 = fmaf(xb, xb, pz * pz);
    bool ring_a = da2 >= lo * lo && da2 <= hi * hi, ring_b = db2 >= lo * lo && db2 <= hi * hi;
    if (rho > rc) {  // yaw only constrains balls that stay clear of the coxa axis
        if (L.wedge != 0.f) {
            // A reachable point has its (own-side) yaw between the limits: it lies on the inner side
            // of both limit lines through the coxa axis, so the ball's centre is at most rc outside
            // of each.  d_min = rho sin(phi - yaw_min), d_max = rho sin(yaw_max - phi); the
            // pi-flipped solution sees the mirrored direction, i.e. both with the opposite sign.
            const float d_min = fmaf(L.cmin, py, -L.smin * px), d_max = fmaf(L.smax, px, -L.cmax * py);
            const float slack = rc + 0.01f;
            ring_a = ring_a && px > -rc && d_min >= -slack && d_max >= -slack;
            ring_b = ring_b && px < rc && d_min <= slack && d_max <= slack;
        } else {
            const float pi = 3.14159265358979f;
            const float phi = atan2f(py, px);
            const float del = asinf(fminf(1.f, rc / rho)) + 1.0e-4f;
            const float phf = phi > 0.f ? phi - pi : phi + pi;  // folded yaw of the flipped solution
            ring_a = ring_a && px > -rc && phi + del >= L.yaw_min && phi - del <= L.yaw_max;
            ring_b = ring_b && px < rc && phf + del >= L.yaw_min && phf - del <= L.yaw_max;
        }
    }
    return ring_a || ring_b;
}

// ---- plane atlas -------------------------------------------------------------------------------
// plane_clamp is a function of the femur-plane point alone, and over most of the plane its outcome
// is "P minus its projection on one particular circle" or "P minus one particular corner" with the
// same winner over a whole neighbourhood.  The atlas is a regular grid over the plane holding, per
// cell, that winner — but only for cells where a Lipschitz bound PROVES every decision of
// plane_clamp (sector tests, validity of P against each circle, arc membership of each projection,
// every pairwise "who is nearer") keeps its sign over the whole cell.  Such a cell is "pure": the
// result for any point in it is a handful of FMAs.  Points in impure cells (or off the atlas) take
// the full evaluation, so the atlas never changes a result, it only skips work.
//
// cell byte: 0 = impure (also what a texture fetch outside the atlas returns); otherwise bit 7 is
// set and bit 6 = P valid, bits 4-5 = sector, bits 0-3 = winner (0-3 circle slot, 4-13 corner
// index + 4; "no candidate at all" is never certified).
struct AtlasView {
    const unsigned char* cells;  // 8 x 4 blocked copy (plain loads)
    cudaTextureObject_t tex;     // the same cells as a 2-D texture (point sampling, border = 0)
    float inv_cell, ox, oy;      // cell coordinates = P * inv_cell + (ox, oy)
    int w, h;
};
constexpr int kAtlasNone = 15;
constexpr unsigned kAtlasPure = 0x80u;
// a cell is certified when every decision margin at its centre exceeds
// kAtlasNeedFactor * cell + kAtlasNeedSlack: the half diagonal (0.7071) widened by 1/128 of a
// cell for the texture unit's fixed-point cell coordinates, plus float rounding of the margins
constexpr float kAtlasNeedFactor = 0.70711f * 1.02f + 0.012f;
constexpr float kAtlasNeedSlack = 2.0e-3f;

struct PlaneProbe {
    int label;           // bits 0-6 as above (never has bit 7)
    float safety;        // every decision keeps its sign within this distance of the probed point
    float valid_safety;  // the same for "P is valid" alone (sector tests + circle margins)
};
// cell byte for a cell that is impure for the distance but whose validity is certified: 1 + valid
constexpr unsigned kAtlasValidOnly = 1u;
LRM_HD unsigned atlas_cell_byte(const PlaneProbe& pr, float need) {
    if (pr.safety > need) return kAtlasPure | (unsigned)pr.label;
    if (pr.valid_safety > need) return kAtlasValidOnly + ((pr.label & 0x40) ? 1u : 0u);
    return 0u;
}

// signed distance of (X, Y) to the decision boundary of an AngleTest, conservatively
LRM_HD float angle_margin(const AngleTest& t, float X, float Y) {
    if (t.c == 0.f && t.ns == 0.f) return 3.0e38f;  // constant outcome
    const float cr = fmaf(t.c, Y, t.ns * X);         // distance to the threshold line
    return fminf(fabsf(cr), fabsf(Y));               // ... or to the X axis, where `up` flips
}

// Winner: the projection on circle (c, r); rival: corner K of the workspace.  plane_clamp lets the
// corner take over when  best_abs^2 > |P - K|^2  (one_leg.cu:109-118).  For a corner ON the circle
// (|K - c| = r), with v = P - c and u = (K - c) / r:
//     |P - K|^2 - (r - |v|)^2  =  2 r s(P),     s(P) = |v| - v . u  >=  0,
// so the corner wins nowhere, and float rounding can only flip the comparison where 2 r s is tiny —
// next to the normal through the corner.  |grad s| = |v / |v| - u| =: chord, and chord changes by at
// most 2 rho / (|v| - rho) within distance rho, hence inside the ball B(P, rho)
//     s  >=  s(P) - rho (chord(P) + 2 rho / m'),        m' = |v| - rho_cap,  rho <= rho_cap.
// Returns the largest such rho (capped) for which 2 r s stays above the rounding noise of both sides
// — (r - |v|)^2 through an approximate rsqrt, |P - K|^2, and the corner being off the circle by a few
// ulp — or 0 if the corner is not on the circle or P is too close to the ray.
LRM_HD float corner_on_circle_safety(float cx, float cy, float r, float kx, float ky, float X, float Y, float ckey) {
    const float ux = kx - cx, uy = ky - cy;
    const float ru = sqrtf(fmaf(ux, ux, uy * uy));
    const float off = fabsf(ru - r);
    if (!(off < 1.0e-3f) || !(r > 1.f)) return 0.f;
    const float vx = X - cx, vy = Y - cy;
    const float m = sqrtf(fmaf(vx, vx, vy * vy));
    const float cap = fminf(0.125f * m, 8.f);
    const float mp = m - cap;
    if (!(mp > 1.f)) return 0.f;
    const float s = m - fmaf(vx, ux, vy * uy) / ru;
    // noise: both squares (relative 4e-7 of (m + r)^2 covers a 3-ulp rsqrt), the corner off the
    // circle, a floor; then as a bound on s, plus the rounding of s itself
    const float tau = 4.0e-7f * (m + r) * (m + r) + 2.f * (ckey + cap) * (off + 2.0e-5f) + 0.02f;
    const float sp = s - (tau / (2.f * r) + 1.0e-6f * m + 1.0e-4f);
    if (!(sp > 0.f)) return 0.f;
    const float chord = sqrtf(2.f * s / m) * 1.001f;
    const float rho = 0.25f * mp * (sqrtf(fmaf(chord, chord, 8.f * sp / mp)) - chord);
    return fminf(cap, 0.98f * rho);
}

// plane_clamp<false> instrumented: same arithmetic, plus the distance within which every decision
// THAT CAN CHANGE THE OUTCOME keeps its sign:
//   * the sector tests that select the circle set (middle, and the saturation test of this side);
//   * "P is valid": when valid, every circle's margin; when invalid, the margin of the most
//     robustly violated circle is enough (the others may flip, the AND stays false);
//   * the winner: its own arc membership, and its lead over every rival — a rival being any corner
//     (when P is invalid) and any circle that is, or within the cell could become, a candidate
//     (a circle that is robustly NOT a candidate needs no lead; one that is far behind needs no
//     robust candidacy).  Every distance involved is 1-Lipschitz, hence the factor 1/2 on leads.
// LABEL = false: only valid / valid_safety are wanted (the reach bits and the choice certificate of
// the volume build): the winner analysis is skipped, `safety` is not meaningful.
template <bool LABEL = true>
LRM_HD PlaneProbe plane_probe(const LegPlan& L, const SectorTable& tab, float X, float Y) {
    const float upf = up_flag(Y);
    const bool upper = angle_gt(L.middle, X, Y, upf);
    float safety = fminf(angle_margin(L.middle, X, Y), angle_margin(L.sat[upper ? 1 : 0], X, Y));
    const int s = find_sector(L, X, Y);
    float cx[4], cy[4], r[4], sg[4], ax[4], ay[4], ah[4];
    cx[0] = 0.f, cy[0] = 0.f, r[0] = L.inner.r, sg[0] = L.inner.sgn;
    const float4 inner_b = tab.row[s][6];
    for (int j = 0; j < 3; j++) {
        const float4 c = tab.row[s][2 * j];
        const float4 a = tab.row[s][2 * j + 1];
        cx[j + 1] = c.x, cy[j + 1] = c.y, r[j + 1] = c.z, sg[j + 1] = c.w;
        ax[j + 1] = a.x, ay[j + 1] = a.y, ah[j + 1] = a.z;
        if (j == 0) ax[0] = a.w;
        if (j == 1) ay[0] = a.w;
        if (j == 2) ah[0] = a.w;
    }
    float key[4], arc[4];
    bool cand[4];
    bool valid = true;
    float valid_margin = 3.0e38f, invalid_margin = 0.f;
    auto arc_margin = [&](float g, float h) {  // g = dot - h*m ; |h| > 1 means constant outcome
        return fabsf(h) > 1.f ? 3.0e38f : 0.5f * fabsf(g);
    };
    for (int j = 0; j < 4; j++) {
        const float vx = X - cx[j], vy = Y - cy[j];
        const float m = sqrtf(fmaf(vx, vx, vy * vy));
        safety = fminf(safety, m - kMarginF);  // the "on the centre" special case
        const float d = r[j] - m;
        key[j] = fabsf(d);
        const float v = sg[j] * d + kMarginF;
        valid = valid & (v > 0.f);
        if (v > 0.f) valid_margin = fminf(valid_margin, v);
        else invalid_margin = fmaxf(invalid_margin, -v);
        const float g = fmaf(vx, ax[j], vy * ay[j]) - ah[j] * m;
        cand[j] = g >= 0.f;
        arc[j] = arc_margin(g, ah[j]);
        if (j == 0) {
            // two arcs: candidate if inside either; robust if robustly inside one or robustly outside both
            const float g2 = fmaf(vx, inner_b.x, vy * inner_b.y) - inner_b.z * m;
            const float arc2 = arc_margin(g2, inner_b.z);
            const bool c2 = g2 >= 0.f;
            if (cand[0] && c2) arc[0] = fmaxf(arc[0], arc2);
            else if (c2) arc[0] = arc2;
            else if (!cand[0]) arc[0] = fminf(arc[0], arc2);
            cand[0] = cand[0] | c2;
        }
    }
    safety = fminf(safety, valid ? valid_margin : invalid_margin);
    const float valid_safety = safety;  // sector tests, circle centres, validity: all reach needs
    if (!LABEL) {
        PlaneProbe early;
        early.label = (valid ? 0x40 : 0) | (s << 4) | kAtlasNone;
        early.safety = 0.f;
        early.valid_safety = valid_safety;
        return early;
    }
    // winner
    int win = kAtlasNone;
    float best = 3.0e38f;
    for (int j = 0; j < 4; j++)
        if (cand[j] && key[j] < best) best = key[j], win = j;
    float ckey[kMaxCorners];
    if (!valid)
        for (int i = 0; i < L.n_corners; i++) {
            const float wx = X - L.corner_x[i], wy = Y - L.corner_y[i];
            ckey[i] = sqrtf(fmaf(wx, wx, wy * wy));
            if (ckey[i] < best) best = ckey[i], win = 4 + i;
        }
    for (int j = 0; j < 4; j++) {
        if (j == win) {
            safety = fminf(safety, arc[j]);
            continue;
        }
        const float lead = 0.5f * (key[j] - best);
        safety = fminf(safety, cand[j] ? lead : fmaxf(arc[j], lead));
    }
    if (!valid)
        for (int i = 0; i < L.n_corners; i++)
            if (4 + i != win) {
                float lead = 0.5f * (ckey[i] - best);
                // A corner that lies ON the winning circle (an end of one of its arcs) can never be
                // nearer than the projection on that circle, but its lead over it only grows with the
                // SQUARE of the distance from the normal through the corner: half the lead certifies
                // nothing for centimetres around that ray.  Certify the comparison itself instead.
                if (win < 4 && lead < 8.f) lead = fmaxf(lead, corner_on_circle_safety(cx[win], cy[win], r[win], L.corner_x[i], L.corner_y[i], X, Y, ckey[i]));
                safety = fminf(safety, lead);
            }
    if (win >= 4 && win != kAtlasNone) safety = fminf(safety, best - 0.01f);  // keep off the corner itself
    if (win == kAtlasNone) safety = fminf(safety, 0.f);  // "nothing qualifies" is never certified
    PlaneProbe out;
    out.label = (valid ? 0x40 : 0) | (s << 4) | win;
    out.safety = safety;
    out.valid_safety = valid_safety;
    return out;
}

// winner table staged next to the sector table: entry [sector*16 + winner] = (cx, cy, r, 0)
struct WinnerTable {
    float4 e[64];
};
LRM_HD void fill_winner_table(const LegPlan& L, WinnerTable* w, int tid, int nthreads) {
    for (int i = tid; i < 64; i += nthreads) {
        const int s = i >> 4, k = i & 15;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k == 0) e = make_float4(0.f, 0.f, L.inner.r, 0.f);
        else if (k < 4) e = make_float4(L.sector[s].slot[k - 1][0], L.sector[s].slot[k - 1][1],
                                        L.sector[s].slot[k - 1][2], 0.f);
        else if (k - 4 < L.n_corners) e = make_float4(L.corner_x[k - 4], L.corner_y[k - 4], 0.f, 0.f);
        w->e[i] = e;
    }
}

// Cells are stored in 8 x 4 blocks (one 32-byte sector each), so the points of a warp, which walk
// a short line through the plane, touch a handful of sectors instead of one per lane.
LRM_HD size_t atlas_index(int w, int ix, int iy) {
    return ((size_t)((iy >> 2) * (w >> 3) + (ix >> 3)) << 5) | (size_t)(((iy & 3) << 3) | (ix & 7));
}
// Cell byte of the cell holding the point with cell coordinates (fx, fy); 0 when it is impure or
// off the atlas.  TEX: one texture instruction does the floor, the bounds check and the blocked
// addressing (unnormalised coordinates, point sampling, border colour 0).
template <bool TEX>
LRM_HD unsigned atlas_fetch(const AtlasView& A, float fx, float fy) {
#ifdef __CUDA_ARCH__
    if (TEX) return tex2D<unsigned char>(A.tex, fx, fy);
#endif
    const int ix = (int)floorf(fx), iy = (int)floorf(fy);
    if ((unsigned)ix >= (unsigned)A.w || (unsigned)iy >= (unsigned)A.h) return 0u;
#ifdef __CUDA_ARCH__
    return __ldg(A.cells + atlas_index(A.w, ix, iy));
#else
    return A.cells[atlas_index(A.w, ix, iy)];
#endif
}
// Plane evaluation of a certified cell: P - (c + r v/|v|), the very operations plane_clamp applies
// to its winner, so that a point gets the same bits whichever tier decides it (a corner is a
// circle of radius 0: c + 0 v = c exactly; certified cells stay off circle centres and corners).
LRM_HD PlaneResult plane_from_label(const WinnerTable& W, unsigned label, float X, float Y) {
    const float4 e = W.e[label & 63];
    const float vx = X - e.x, vy = Y - e.y;
    const float k = e.z * fast_rsqrt(fmaf(vx, vx, vy * vy));
    PlaneResult out;
    out.valid = (label & 0x40u) != 0;
    out.dx = X - fmaf(vx, k, e.x);
    out.dy = Y - fmaf(vy, k, e.y);
    return out;
}

struct CoxaPoint {
    float x, y, z;  // point in the coxa frame
};

LRM_HD CoxaPoint to_coxa_frame(const LegPlan& L, float x, float y, float z) {
    CoxaPoint p;
    p.x = fmaf(L.M[0], x, fmaf(L.M[1], y, fmaf(L.M[2], z, L.t[0])));
    p.y = fmaf(L.M[3], x, fmaf(L.M[4], y, fmaf(L.M[5], z, L.t[1])));
    p.z = fmaf(L.M[6], x, fmaf(L.M[7], y, fmaf(L.M[8], z, L.t[2])));
    return p;
}

// reachability_circles, one_leg.cu:280-319
LRM_HD bool reach_coxa_frame(const LegPlan& L, const SectorTable& tab, const CoxaPoint p) {
    const bool flip = f2i(p.x) < 0;  // signbit: mirrored through the coxa axis
    const float xf = flip ? -p.x : p.x;
    const float yf = flip ? -p.y : p.y;
    if (angle_gt(L.over, xf, yf) | angle_gt(L.under, xf, -yf)) return false;  // outside the yaw limits
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float rho = rho2 > 0.f ? rho2 * fast_rsqrt(rho2) : 0.f;
    const float X = (flip ? -rho : rho) - L.coxa_length;
    return plane_reach(L, tab, X, p.z);
}

// reachability_circles through the plane atlas: the coxa yaw tests are evaluated as always, the
// four circle tests of eval_plane_circles<REACH> come from the certified valid bit of the cell
// holding the plane point; cells certified neither for the distance nor for validity alone
// (a band around the reachability edge and the sector lines) take the explicit tests.
template <bool TEX>
LRM_HD bool reach_coxa_frame_atlas(const LegPlan& L, const SectorTable& tab, const AtlasView& A,
                                   const CoxaPoint p) {
    const bool flip = f2i(p.x) < 0;  // signbit: mirrored through the coxa axis
    const float xf = flip ? -p.x : p.x;
    const float yf = flip ? -p.y : p.y;
    if (angle_gt(L.over, xf, yf) | angle_gt(L.under, xf, -yf)) return false;  // outside the yaw limits
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float rho = rho2 > 0.f ? rho2 * fast_rsqrt(rho2) : 0.f;
    const float X = (flip ? -rho : rho) - L.coxa_length;
    const unsigned cell = atlas_fetch<TEX>(A, fmaf(X, A.inv_cell, A.ox), fmaf(p.z, A.inv_cell, A.oy));
    if (cell & kAtlasPure) return (cell & 0x40u) != 0;
    if (cell != 0u) return cell != kAtlasValidOnly;
    return plane_reach(L, tab, X, p.z);
}

// The yaw tests of finish_finding_closest (one_leg.cu:222-234) for the solution whose yaw is the
// angle of (wx, wy).
struct YawFlags {
    bool mega, over, under, upper_lim;
};

// Both coxa solutions at once.  The direct one has yaw = angle of (x, y), the flipped one
// yaw -+ pi = angle of (-x, 0 - y) ("0 - y" keeps atan2f's +pi, not -pi, for y = +0, x > 0, like
// coxangle + pi does in the reference, one_leg.cu:329).  For a test with cross product
// ip = c*Y + ns*X the flipped solution sees -ip, so each inner product is shared:
//   direct:  ip + bias > thr(up)      <=>  ip >  thr(up_d) - bias
//   flipped: bias - ip > thr(up)      <=>  ip <  bias - thr(up_f)
LRM_HD void yaw_pair(const AngleTest& t, float ip, float up_d, float up_f, bool& d, bool& f) {
    const float k = t.thr_dn - t.bias;  // uniform
    d = ip > fmaf(up_d, -kAngleBig, k);
    f = ip < fmaf(up_f, kAngleBig, -k);
}
LRM_HD void yaw_tests_both(const LegPlan& L, float x, float y, YawFlags& a, YawFlags& b) {
    const float yf = 0.f - y;
    // "greater" tests look at (X, Y), "less" tests at (X, -Y)
    const float up_d = up_flag(y), up_dn = up_flag(-y);
    const float up_f = up_flag(yf), up_fn = up_flag(-yf);
    bool hi_d, hi_f, lo_d, lo_f;
    yaw_pair(L.mega_hi, fmaf(L.mega_hi.c, y, L.mega_hi.ns * x), up_d, up_f, hi_d, hi_f);
    yaw_pair(L.mega_lo, fmaf(L.mega_lo.c, -y, L.mega_lo.ns * x), up_dn, up_fn, lo_d, lo_f);
    a.mega = hi_d | lo_d, b.mega = hi_f | lo_f;
    yaw_pair(L.over, fmaf(L.over.c, y, L.over.ns * x), up_d, up_f, a.over, b.over);
    yaw_pair(L.under, fmaf(L.under.c, -y, L.under.ns * x), up_dn, up_fn, a.under, b.under);
    yaw_pair(L.mid, fmaf(L.mid.c, y, L.mid.ns * x), up_d, up_f, a.upper_lim, b.upper_lim);
}

// The yaw decisions of both solutions of a direction as one small integer: solution kind
// [free lo, free hi, min lo, min hi, max lo, max hi, mega] ("hi" = upper_lim), kYawSkipped for a
// solution that duplicates the other one (same rule as dist_coxa_frame); direct kind in bits 0-2,
// flipped kind in bits 3-5.
constexpr int kYawSkipped = 7;
LRM_HD int yaw_sol_kind(const YawFlags& f) {
    if (f.mega) return 6;
    if (f.under) return 2 + (f.upper_lim ? 1 : 0);
    if (f.over) return 4 + (f.upper_lim ? 1 : 0);
    return f.upper_lim ? 1 : 0;
}
LRM_HD int yaw_combo(const LegPlan& L, float x, float y) {
    YawFlags fa, fb;
    yaw_tests_both(L, x, y, fa, fb);
    const bool skip_a = fa.mega & !(fb.mega | fb.over | fb.under);
    const bool skip_b = fb.mega & !(fa.mega | fa.over | fa.under);
    return ((skip_b ? kYawSkipped : yaw_sol_kind(fb)) << 3) | (skip_a ? kYawSkipped : yaw_sol_kind(fa));
}

struct BranchResult {
    bool res;          // was_valid && !coxa_saturated
    float vx, vy, vz;  // vector in the coxa frame
    float n2;          // its squared norm
};

// finish_finding_closest<bool>, one_leg.cu:215-278, for one coxa solution, in two halves around
// the plane evaluation.  (ux, uy) = unit vector of the solution's un-saturated yaw (w / rho).
struct BranchPrep {
    float cs, ss;  // unit direction of the saturated yaw
    float X, yr;   // femur-plane abscissa of the point and its out-of-plane offset
    bool saturated;
};
LRM_HD BranchPrep branch_prep(const LegPlan& L, const CoxaPoint p, const YawFlags f, float ux,
                              float uy) {
    BranchPrep b;
    b.cs = ux, b.ss = uy;
    if (f.mega) {
        b.cs = -ux, b.ss = -uy;  // yaw -+ pi
    } else if (f.under) {
        b.cs = L.cos_min, b.ss = L.sin_min;
    } else if (f.over) {
        b.cs = L.cos_max, b.ss = L.sin_max;
    }
    b.saturated = f.mega | f.over | f.under;
    b.X = fmaf(p.x, b.cs, p.y * b.ss) - L.coxa_length;
    b.yr = fmaf(p.y, b.cs, -p.x * b.ss);
    return b;
}
LRM_HD BranchResult branch_finish(const LegPlan& L, const CoxaPoint p, const YawFlags f,
                                  const BranchPrep b, const PlaneResult pl) {
    const float qx = pl.dx, qy = b.yr, qz = pl.dy;  // in the saturated-yaw frame
    const float n2 = fmaf(qx, qx, fmaf(qy, qy, qz * qz));
    BranchResult out;
    out.res = pl.valid & !b.saturated;
    // in-plane region reached but a coxa-limit half-plane is nearer (one_leg.cu:258-274)
    const float cl = f.upper_lim ? L.cos_max : L.cos_min;
    const float sl = f.upper_lim ? L.sin_max : L.sin_min;
    const float yl = fmaf(p.y, cl, -p.x * sl);
    const bool to_plane = pl.valid & !f.mega & (n2 > yl * yl);
    out.vx = to_plane ? -yl * sl : fmaf(qx, b.cs, -qy * b.ss);
    out.vy = to_plane ? yl * cl : fmaf(qx, b.ss, qy * b.cs);
    out.vz = to_plane ? 0.f : qz;
    out.n2 = to_plane ? yl * yl : n2;
    return out;
}

struct DistResult {
    bool flag;        // distance_circles' return: res || resflip
    bool reach;       // reachability_circles of the same point
    float dx, dy, dz; // world-frame vector
};

// distance_circles (one_leg.cu:321-341) + the way back to the world frame: the full evaluation.
template <bool GENERIC>
LRM_HD DistResult dist_coxa_frame(const LegPlan& L, const SectorTable& tab, const CoxaPoint p) {
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float inv_rho = rho2 > 0.f ? fast_rsqrt(rho2) : 0.f;
    // unit vector of the direct yaw; a point on the coxa axis has yaw 0 (atan2f(0, 0))
    const float ux = rho2 > 0.f ? p.x * inv_rho : 1.f;
    const float uy = rho2 > 0.f ? p.y * inv_rho : 0.f;
    YawFlags fa, fb;
    yaw_tests_both(L, p.x, p.y, fa, fb);
    // A solution beyond limit +- pi/2 is evaluated in the plane of the OTHER solution's yaw
    // (one_leg.cu:225-226).  When that other solution is unsaturated both plane evaluations are
    // identical and the mega one can never be preferred (it reports res = false and the same
    // vector unless the other one found something nearer), so it is skipped.
    const bool skip_a = fa.mega & !(fb.mega | fb.over | fb.under);
    const bool skip_b = fb.mega & !(fa.mega | fa.over | fa.under);
    const BranchPrep pa = branch_prep(L, p, fa, ux, uy);
    const BranchPrep pb = branch_prep(L, p, fb, -ux, -uy);
    BranchResult a, b;
    a.res = b.res = false, a.vx = a.vy = a.vz = b.vx = b.vy = b.vz = 0.f, a.n2 = b.n2 = 0.f;
    if (!skip_a) a = branch_finish(L, p, fa, pa, plane_clamp<GENERIC>(L, tab, pa.X, p.z));
    if (!skip_b) b = branch_finish(L, p, fb, pb, plane_clamp<GENERIC>(L, tab, pb.X, p.z));
    if (skip_a) a = b, a.res = false;
    if (skip_b) b = a, b.res = false;
    const bool direct = (a.res == b.res) ? (a.n2 < b.n2) : a.res;
    const float vx = direct ? a.vx : b.vx, vy = direct ? a.vy : b.vy, vz = direct ? a.vz : b.vz;
    DistResult out;
    out.flag = a.res | b.res;
    out.reach = (f2i(p.x) < 0) ? b.res : a.res;
    out.dx = fmaf(L.Mo[0], vx, fmaf(L.Mo[1], vy, L.Mo[2] * vz));
    out.dy = fmaf(L.Mo[3], vx, fmaf(L.Mo[4], vy, L.Mo[5] * vz));
    out.dz = fmaf(L.Mo[6], vx, fmaf(L.Mo[7], vy, L.Mo[8] * vz));
    return out;
}

// ---- fast path: yaw-sector table + plane atlas --------------------------------------------------
// Same arithmetic as dist_coxa_frame for everything that produces a number (plane direction,
// plane abscissa, projection on the winner, limit-plane rule, final choice); the DECISIONS come
// from two certified tables: the yaw-sector code of the point's direction (FastTables) and the
// plane-atlas cell of each solution's plane point.  Returns false — nothing is written — when
// either table cannot certify the point; the caller then runs dist_coxa_frame.
struct FastView {
    const YawPair* pair;        // [kYawPairs]
    const unsigned char* code;  // [kYawBins + 1]
    const int32_t* combo;       // [ncombo]: the yaw_combo each pair stands for (EXACT_YAW only)
    int ncombo;
};

// Diamond angle bin of a unit direction: d = uy / (|ux| + |uy|) in the right half plane, mirrored
// to +-(2 - |d|) in the left one: monotone in atan2(uy, ux), -2 at -pi, +2 at +pi.
LRM_HD int yaw_bin(float ux, float uy) {
    const float d0 = uy * fast_rcp(fabsf(ux) + fabsf(uy));
    const float d = f2i(ux) < 0 ? copysignf(2.f, uy) - d0 : d0;
    return (int)fmaf(d, 0.25f * kYawBins, 0.5f * kYawBins);
}

LRM_HD BranchResult fast_branch(const CoxaPoint p, const YawSol& s, float cs, float ss, float yr,
                                float yl, const PlaneResult pl) {
    const float qx = pl.dx, qy = yr, qz = pl.dy;  // in the plane's frame
    const float n2 = fmaf(qx, qx, fmaf(qy, qy, qz * qz));
    // in-plane region reached but the coxa-limit half-plane is nearer (one_leg.cu:258-274);
    // s.big = +inf switches the rule off for a mega-saturated yaw
    const float yl2 = fmaf(yl, yl, s.big);
    const bool to_plane = pl.valid & (n2 > yl2);
    BranchResult out;
    out.res = pl.valid & (s.nsat != 0.f);
    out.vx = to_plane ? -yl * s.sl : fmaf(qx, cs, -qy * ss);
    out.vy = to_plane ? yl * s.cl : fmaf(qx, ss, qy * cs);
    out.vz = to_plane ? 0.f : qz;
    out.n2 = to_plane ? yl2 : n2;
    return out;
}

// Straight-line on purpose (one exit, no early returns): a thread's two consecutive points can
// then be interleaved by the compiler, so that their texture fetches overlap.  A point the tables
// cannot certify runs to the end on harmless values and reports false.
// EXACT_YAW (the dense redo of the tiered sweep): a direction whose bin is uncertified — it holds
// a decision boundary — has its yaw decisions evaluated for the point itself (yaw_combo: the very
// tests the full evaluation runs) and mapped to its solution pair, instead of giving up.
template <bool TEX, bool SKIP_B = true, bool EXACT_YAW = false>
LRM_HD bool dist_fast(const LegPlan& L, const FastView& F, const AtlasView& A, const WinnerTable& W,
                      const CoxaPoint p, DistResult* out) {
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    bool ok = rho2 > 1.0e-12f;  // on the coxa axis (or NaN): full evaluation
    const float inv_rho = fast_rsqrt(ok ? rho2 : 1.f);
    const float ux = p.x * inv_rho, uy = p.y * inv_rho;
    int bin = yaw_bin(ux, uy);
    ok = ok & ((unsigned)bin <= (unsigned)kYawBins);
    bin = ok ? bin : 0;
    unsigned code = F.code[bin];
    if (EXACT_YAW) {
        if (ok && code == kYawImpure) {
            const int combo = yaw_combo(L, p.x, p.y);
            for (int i = 0; i < F.ncombo; i++)
                if (F.combo[i] == combo) code = (unsigned)i;
        }
    }
    ok = ok & (code != kYawImpure);
    code = ok ? code : 0u;
    const YawPair& pr = F.pair[code];
    const YawSol& sa = pr.a;
    const YawSol& sb = pr.b;
    const bool has_a = sa.present != 0.f, has_b = sb.present != 0.f;
    const float csa = fmaf(sa.k, ux, sa.c_cs), ssa = fmaf(sa.k, uy, sa.c_ss);
    const float csb = fmaf(sb.k, ux, sb.c_cs), ssb = fmaf(sb.k, uy, sb.c_ss);
    const float Xa = fmaf(p.x, csa, p.y * ssa) - L.coxa_length;
    const float Xb = fmaf(p.x, csb, p.y * ssb) - L.coxa_length;
    // both cells are requested before either is consumed; an absent solution reads its cell too
    // (its plane point is a valid one) and is neutralised below
    const float fy = fmaf(p.z, A.inv_cell, A.oy);
    const unsigned la = atlas_fetch<TEX>(A, fmaf(Xa, A.inv_cell, A.ox), fy);
    const unsigned lb = atlas_fetch<TEX>(A, fmaf(Xb, A.inv_cell, A.ox), fy);
    ok = ok & ((((has_a ? la : kAtlasPure) & (has_b ? lb : kAtlasPure)) & kAtlasPure) != 0u);
    const float yra = fmaf(p.y, csa, -p.x * ssa), yla = fmaf(p.y, sa.cl, -p.x * sa.sl);
    const float yrb = fmaf(p.y, csb, -p.x * ssb), ylb = fmaf(p.y, sb.cl, -p.x * sb.sl);
    // a solution that is absent can never be preferred: res = false, infinitely far
    BranchResult a = fast_branch(p, sa, csa, ssa, yra, yla, plane_from_label(W, la, Xa, p.z));
    a.res = a.res & has_a;
    a.n2 = has_a ? a.n2 : 3.0e38f;
    BranchResult b;
    b.res = false, b.vx = b.vy = b.vz = 0.f, b.n2 = 3.0e38f;
    if (SKIP_B) {
        // A saturated flipped solution reports res = false and a vector at least as long as its
        // offset from its own plane (or from its limit plane): when the direct one is valid, or
        // already nearer than that, the flipped one cannot be chosen and is not evaluated.
        const float floor_b = fminf(yrb * yrb, fmaf(ylb, ylb, sb.big));
        const bool need_b = has_b & !((sb.nsat == 0.f) & has_a & (a.res | (a.n2 < floor_b)));
        if (need_b) b = fast_branch(p, sb, csb, ssb, yrb, ylb, plane_from_label(W, lb, Xb, p.z));
    } else {
        b = fast_branch(p, sb, csb, ssb, yrb, ylb, plane_from_label(W, lb, Xb, p.z));
        b.res = b.res & has_b;
        b.n2 = has_b ? b.n2 : 3.0e38f;
    }
    const bool direct = (a.res == b.res) ? (a.n2 < b.n2) : a.res;
    const float vx = direct ? a.vx : b.vx, vy = direct ? a.vy : b.vy, vz = direct ? a.vz : b.vz;
    out->flag = a.res | b.res;
    out->reach = (f2i(p.x) < 0) ? b.res : a.res;
    out->dx = fmaf(L.Mo[0], vx, fmaf(L.Mo[1], vy, L.Mo[2] * vz));
    out->dy = fmaf(L.Mo[3], vx, fmaf(L.Mo[4], vy, L.Mo[5] * vz));
    out->dz = fmaf(L.Mo[6], vx, fmaf(L.Mo[7], vy, L.Mo[8] * vz));
    return ok;
}


// ---- choice volume: which coxa solution wins, certified per 3-D cell ----------------------------
// dist_fast still evaluates BOTH coxa solutions and needs BOTH plane cells certified.  Which of
// the two wins (one_leg.cu:334), and every yaw decision behind them, varies slowly in space: a
// regular grid of cubes over the coxa frame holds, per cube, the winning solution — index of its
// YawSol in the pair table — wherever that can be PROVEN constant over the cube:
//   * the yaw decisions: same yaw_combo at the four corners of the cube's (padded) xy footprint,
//     footprint clear of the y = 0 plane (signed-zero rules of atan2f, the +-pi seam).  The
//     decision boundaries are rays from the coxa axis, the footprint is convex, so a boundary
//     crossing it separates two of its corners;
//   * the choice: the length of each solution's vector is 1-Lipschitz in the point (distance to
//     a closed set in an isometric frame), so |n_other - n_chosen| / 2 bounds the radius within
//     which the comparison keeps its sign; a solution with an unsaturated yaw additionally
//     flips `res` where its plane point crosses the reachability edge, bounded by the plane
//     probe's valid_safety.  Checked at the cube centre against the cube's half diagonal, and
//     if that fails on 4^3 sub-cubes against theirs.
// A point in a certified cube evaluates ONE solution: one plane-atlas fetch, one projection.  The
// limit-plane rule (one_leg.cu:258-274) is still evaluated per point.  Uncertified cubes (and
// everything off the volume) take dist_fast, uncertified plane cells the full evaluation.
//
// cube byte: 0 = uncertified; else bit 7 | pair index << 1 | side (0 direct, 1 flipped) — bits 0-4
// index the pair table viewed as YawSol[2 * kYawPairs].
constexpr unsigned kVolPure = 0x80u;
// bits 5-6, independent of the choice bits: reachability_circles is the same for every point of the
// cube (bit 5), and its value (bit 6) — the reach-only sweep reads nothing else for such a point
constexpr unsigned kVolReachKnown = 0x20u, kVolReachValue = 0x40u;
struct VolumeView {
    cudaTextureObject_t tex;  // 3-D texture of 32-bit cube texels (point sampling, border = 0)
    float inv_cell, o, oy;    // cube coordinates = p * inv_cell + o (x, z), + oy (y)
    int dim;
    const unsigned short* bricks;  // fine texels of the bricks, 64 per brick (x fastest); nullptr: none
};
// The grid is shifted by half a cube along y: the y = 0 plane (a decision boundary for every leg:
// the sign of y picks the coxa limit of the limit-plane rule, and carries atan2f's signed-zero
// rules) then cuts ONE layer of cubes through the middle instead of touching two.
constexpr float kVolShiftY = 0.5f;

// Both solutions of one point through the full evaluation, with what the certification needs.
struct ChoiceProbe {
    bool direct;   // the full evaluation's choice at this point
    float margin;  // the choice (and the flags that follow from it) cannot change within this distance
};
LRM_HD ChoiceProbe choice_probe(const LegPlan& L, const SectorTable& tab, const CoxaPoint p) {
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float inv_rho = rho2 > 0.f ? fast_rsqrt(rho2) : 0.f;
    const float ux = rho2 > 0.f ? p.x * inv_rho : 1.f;
    const float uy = rho2 > 0.f ? p.y * inv_rho : 0.f;
    YawFlags fa, fb;
    yaw_tests_both(L, p.x, p.y, fa, fb);
    const BranchPrep pa = branch_prep(L, p, fa, ux, uy);
    const BranchPrep pb = branch_prep(L, p, fb, -ux, -uy);
    const PlaneResult pla = plane_clamp<false>(L, tab, pa.X, p.z);
    const PlaneResult plb = plane_clamp<false>(L, tab, pb.X, p.z);
    const BranchResult a = branch_finish(L, p, fa, pa, pla);
    const BranchResult b = branch_finish(L, p, fb, pb, plb);
    ChoiceProbe out;
    out.direct = (a.res == b.res) ? (a.n2 < b.n2) : a.res;
    const float nc = sqrtf(out.direct ? a.n2 : b.n2), no = sqrtf(out.direct ? b.n2 : a.n2);
    const bool c_sat = out.direct ? pa.saturated : pb.saturated, o_sat = out.direct ? pb.saturated : pa.saturated;
    const bool c_valid = out.direct ? pla.valid : plb.valid, o_valid = out.direct ? plb.valid : pla.valid;
    const float Xc = out.direct ? pa.X : pb.X, Xo = out.direct ? pb.X : pa.X;
    // the chosen one stays strictly nearer ...  (0.05 mm: float rounding of both lengths)
    float m = 0.5f * (no - nc) - 0.05f;
    // ... or stays valid with an unsaturated yaw: res = true decides for it whatever the lengths
    if (!c_sat && c_valid) m = fmaxf(m, plane_probe<false>(L, tab, Xc, p.z).valid_safety - 0.01f);
    // the other one must not turn res = true anywhere nearby
    if (!o_sat) m = o_valid ? -1.f : fminf(m, plane_probe<false>(L, tab, Xo, p.z).valid_safety - 0.01f);
    out.margin = m;
    return out;
}

// Cube byte of the cube [x0, x0+h] x [y0, y0+h] x [z0, z0+h] of the coxa frame, in two steps so
// that the device build can share the refinement of a cube among the lanes of a warp:
// choice_cell_first decides everything the cube centre can decide, choice_cell_sub checks one of
// the 4^3 sub-cubes.  choice_cell_byte is the serial composition (host emulation, reference).
constexpr int kVolSub = 4;  // sub-cubes per axis of the refinement
struct CellFirst {
    unsigned byte;  // the choice bits if certified (after refinement, when refine is set), else 0
    bool refine;    // the centre could not decide: all kVolSub^3 sub-cubes must pass choice_cell_sub
    bool direct;    // the centre's choice (what every sub-cube must agree with)
    unsigned reach;      // the reach bits if certified (after refinement, when reach_refine is set), else 0
    bool reach_refine;   // ... all sub-cubes must pass reach_cell_sub
    bool reach_flip;     // the cube lies behind the coxa axis (x < 0): the pi-flipped solution's plane
};
LRM_HD float vol_pad(float h) {
    // the texture unit resolves cube coordinates to 1/256 of a cube; float rounding of the
    // coordinate itself is far below 0.01 mm
    return h * (1.f / 64.f) + 0.01f;
}
// pad_h (all cube functions): the cube size whose texture-coordinate slack applies — the cube's own
// (0: a cube of the coarse grid), or the coarse cube's for a fine cube of a brick (the sweep finds the
// brick through the coarse texel fetch).
LRM_HD CellFirst choice_cell_first(const LegPlan& L, const SectorTable& tab, const FastTables& FT, float x0,
                                   float y0, float z0, float h, float pad_h = 0.f) {
    CellFirst out;
    out.byte = 0u, out.refine = false, out.direct = true;
    out.reach = 0u, out.reach_refine = false, out.reach_flip = false;
    const float pad = vol_pad(pad_h > 0.f ? pad_h : h);
    const float xa = x0 - pad, xb = x0 + h + pad, ya = y0 - pad, yb = y0 + h + pad;
    if (ya <= 0.f && yb >= 0.f) return out;
    const int combo = yaw_combo(L, xa, ya);
    if (yaw_combo(L, xb, ya) != combo || yaw_combo(L, xa, yb) != combo || yaw_combo(L, xb, yb) != combo)
        return out;
    // reachability_circles (one_leg.cu:280-319): the solution on the point's own side of the coxa
    // axis, unsaturated yaw, plane point valid.  Needs the cube clear of the x = 0 plane too.  The
    // plane point (+-rho - coxa, z) moves at most as far as the point does, so plane_probe's
    // valid_safety at the centre bounds the radius within which validity cannot change.
    if (xa > 0.f || xb < 0.f) {
        out.reach_flip = xb < 0.f;
        const int kind = out.reach_flip ? (combo >> 3) : (combo & 7);
        if (kind > 1) {
            out.reach = kVolReachKnown;  // yaw outside the coxa limits all over the cube: unreachable
        } else {
            const float side = h + 2.f * pad, hs = 0.5f * side;
            const float cx = xa + hs, cy = ya + hs, cz = z0 - pad + hs;
            const float rho = sqrtf(fmaf(cx, cx, cy * cy));
            const PlaneProbe pr = plane_probe<false>(L, tab, (out.reach_flip ? -rho : rho) - L.coxa_length, cz);
            const unsigned bits = kVolReachKnown | ((pr.label & 0x40) ? kVolReachValue : 0u);
            if (pr.valid_safety > 0.8660255f * side + 0.01f) {
                out.reach = bits;
            } else if (pr.valid_safety > 0.8660255f * side * (1.f / kVolSub) + 0.01f) {
                out.reach = bits, out.reach_refine = true;
            }
        }
    }
    int id = -1;
    for (int i = 0; i < FT.ncombo; i++)
        if (FT.combo[i] == combo) id = i;
    if (id < 0 || FT.both_unsat) return out;
    const bool has_a = (combo & 7) != kYawSkipped, has_b = (combo >> 3) != kYawSkipped;
    if (!has_b || !has_a) {
        out.byte = kVolPure | ((unsigned)id << 1) | (has_b ? 1u : 0u);
        return out;
    }
    const float side = h + 2.f * pad;
    const float r0 = 0.8660255f * side;
    CoxaPoint c;
    c.x = xa + 0.5f * side, c.y = ya + 0.5f * side, c.z = z0 - pad + 0.5f * side;
    const ChoiceProbe pc = choice_probe(L, tab, c);
    out.direct = pc.direct;
    if (pc.margin > r0) {
        out.byte = kVolPure | ((unsigned)id << 1) | (pc.direct ? 0u : 1u);
    } else if (pc.margin > 0.25f * r0) {
        // Refinement needs every sub-cube centre to clear r0 / 4.  Below -r0 / 2 at the centre that is
        // impossible (the margin is 1-Lipschitz), below r0 / 4 it would take a local minimum of the
        // margin at the centre: not worth 64 probes.
        out.byte = kVolPure | ((unsigned)id << 1) | (pc.direct ? 0u : 1u);
        out.refine = true;
    }
    return out;
}
LRM_HD bool choice_cell_sub(const LegPlan& L, const SectorTable& tab, float x0, float y0, float z0, float h,
                            int k, bool direct, float pad_h = 0.f) {
    const float pad = vol_pad(pad_h > 0.f ? pad_h : h);
    const float side = h + 2.f * pad;
    const float sub = side * (1.f / kVolSub), r1 = 0.8660255f * sub;
    CoxaPoint q;
    q.x = x0 - pad + ((float)(k % kVolSub) + 0.5f) * sub;
    q.y = y0 - pad + ((float)((k / kVolSub) % kVolSub) + 0.5f) * sub;
    q.z = z0 - pad + ((float)(k / (kVolSub * kVolSub)) + 0.5f) * sub;
    const ChoiceProbe ps = choice_probe(L, tab, q);
    return ps.direct == direct && ps.margin > r1;
}
// sub-cube k of a cube whose centre left the reach bit open: same validity, safely
LRM_HD bool reach_cell_sub(const LegPlan& L, const SectorTable& tab, float x0, float y0, float z0, float h,
                           int k, bool flip, bool valid, float pad_h = 0.f) {
    const float pad = vol_pad(pad_h > 0.f ? pad_h : h);
    const float side = h + 2.f * pad;
    const float sub = side * (1.f / kVolSub), r1 = 0.8660255f * sub;
    const float qx = x0 - pad + ((float)(k % kVolSub) + 0.5f) * sub;
    const float qy = y0 - pad + ((float)((k / kVolSub) % kVolSub) + 0.5f) * sub;
    const float qz = z0 - pad + ((float)(k / (kVolSub * kVolSub)) + 0.5f) * sub;
    const float rho = sqrtf(fmaf(qx, qx, qy * qy));
    const PlaneProbe pr = plane_probe<false>(L, tab, (flip ? -rho : rho) - L.coxa_length, qz);
    return ((pr.label & 0x40) != 0) == valid && pr.valid_safety > r1 + 0.01f;
}
LRM_HD unsigned choice_cell_byte(const LegPlan& L, const SectorTable& tab, const FastTables& FT, float x0,
                                 float y0, float z0, float h, float pad_h = 0.f) {
    const CellFirst f = choice_cell_first(L, tab, FT, x0, y0, z0, h, pad_h);
    unsigned byte = f.byte, reach = f.reach;
    if (f.refine)
        for (int k = 0; k < kVolSub * kVolSub * kVolSub && byte; k++)
            if (!choice_cell_sub(L, tab, x0, y0, z0, h, k, f.direct, pad_h)) byte = 0u;
    if (f.reach_refine)
        for (int k = 0; k < kVolSub * kVolSub * kVolSub && reach; k++)
            if (!reach_cell_sub(L, tab, x0, y0, z0, h, k, f.reach_flip, (reach & kVolReachValue) != 0, pad_h)) reach = 0u;
    return byte | reach;
}

// ---- plane label of a cube: tier 0 of the distance sweep ----------------------------------------
// dist_choice still needs the plane-atlas cell of the chosen solution's plane point: a second
// fetch that DEPENDS on the first.  The plane point (X, z) of a solution is a 1-Lipschitz image of
// the point (X = +-rho - coxa for an unsaturated / mega yaw, a fixed direction's abscissa at a coxa
// limit), so over a cube it stays inside a small rectangle of the plane: where the whole
// rectangle carries ONE certified atlas label, that label is a property of the cube and travels in
// the high byte of the cube's texel — one fetch per point, no dependent one.
// texel (16 bit): low byte = cube byte as above; high byte = 0, or the atlas cell byte (bit 7 set,
// bit 6 valid, bits 4-5 sector, bits 0-3 winner) shared by every plane point of the cube.
//
// Plane abscissa of one solution of p, exactly as dist_coxa_frame builds it (branch_prep).
LRM_HD float solution_plane_x(const LegPlan& L, const CoxaPoint p, bool flipped) {
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float inv_rho = rho2 > 0.f ? fast_rsqrt(rho2) : 0.f;
    const float ux = rho2 > 0.f ? p.x * inv_rho : 1.f;
    const float uy = rho2 > 0.f ? p.y * inv_rho : 0.f;
    YawFlags fa, fb;
    yaw_tests_both(L, p.x, p.y, fa, fb);
    return flipped ? branch_prep(L, p, fb, -ux, -uy).X : branch_prep(L, p, fa, ux, uy).X;
}
// slack on the rectangle: float rounding of the plane point in the sweep vs here (< 1e-3 mm)
constexpr float kPlaneRectSlack = 0.02f;
// (a) by the instrumented evaluation at the centre: every decision keeps its sign within the
//     cube's half diagonal.  Used on blocks of cubes (one probe settles 64 cubes of the far field).
LRM_HD unsigned cube_plane_probe(const LegPlan& L, const SectorTable& tab, float Xc, float zc, float side) {
    const PlaneProbe pr = plane_probe(L, tab, Xc, zc);
    return pr.safety > 0.8660255f * side + kPlaneRectSlack ? (kAtlasPure | (unsigned)pr.label) : 0u;
}
// (b) by the atlas itself: every cell meeting the rectangle [Xc +- d_xy] x [zc +- side / 2] is
//     certified with the same label (d_xy = half diagonal of the xy footprint: |dX| <= |d(x, y)|).
//     Each certified cell vouches for all of its points, so the union vouches for the cube.
LRM_HD unsigned cube_plane_scan(const AtlasView& A, float Xc, float zc, float side) {
    const float dx = 0.70710678f * side + kPlaneRectSlack, dz = 0.5f * side + kPlaneRectSlack;
    const int ix0 = (int)floorf(fmaf(Xc - dx, A.inv_cell, A.ox)), ix1 = (int)floorf(fmaf(Xc + dx, A.inv_cell, A.ox));
    const int iy0 = (int)floorf(fmaf(zc - dz, A.inv_cell, A.oy)), iy1 = (int)floorf(fmaf(zc + dz, A.inv_cell, A.oy));
    if (ix0 < 0 || iy0 < 0 || ix1 >= A.w || iy1 >= A.h) return 0u;
    const unsigned ref = A.cells[atlas_index(A.w, ix0, iy0)];
    if (!(ref & kAtlasPure)) return 0u;
    for (int iy = iy0; iy <= iy1; iy++)
        for (int ix = ix0; ix <= ix1; ix++)
            if (A.cells[atlas_index(A.w, ix, iy)] != ref) return 0u;
    return ref;
}
// Centre of the padded cube and its side, as choice_cell_first sees them.
LRM_HD CoxaPoint cube_centre(float x0, float y0, float z0, float h, float* side, float pad_h = 0.f) {
    const float pad = vol_pad(pad_h > 0.f ? pad_h : h);
    *side = h + 2.f * pad;
    CoxaPoint c;
    c.x = x0 - pad + 0.5f * *side, c.y = y0 - pad + 0.5f * *side, c.z = z0 - pad + 0.5f * *side;
    return c;
}
// Texel shared by a whole block of cubes (the coarse pass of the volume build), or 0 when the block
// is not settled as a whole: the centre must decide the choice, the reach bits AND the plane label
// without any refinement.
LRM_HD unsigned coarse_block_word(const LegPlan& L, const SectorTable& tab, const FastTables& FT, float x0, float y0,
                                  float z0, float h) {
    const CellFirst f = choice_cell_first(L, tab, FT, x0, y0, z0, h);
    if (f.byte == 0u || f.refine || f.reach == 0u || f.reach_refine) return 0u;
    float side;
    const CoxaPoint c = cube_centre(x0, y0, z0, h, &side);
    const unsigned hi = cube_plane_probe(L, tab, solution_plane_x(L, c, (f.byte & 1u) != 0u), c.z, side);
    return hi ? (f.byte | f.reach | (hi << 8)) : 0u;
}
// Whole texel of one cube, serially (host emulation; the device build splits the work):
// probe = true certifies the plane label as the coarse pass does, false as the fine pass does.
LRM_HD unsigned choice_cell_word(const LegPlan& L, const SectorTable& tab, const FastTables& FT, const AtlasView& A,
                                 float x0, float y0, float z0, float h, bool probe, float pad_h = 0.f) {
    const unsigned lo = choice_cell_byte(L, tab, FT, x0, y0, z0, h, pad_h);
    if (!(lo & kVolPure)) return lo;
    float side;
    const CoxaPoint c = cube_centre(x0, y0, z0, h, &side, pad_h);
    const float Xc = solution_plane_x(L, c, (lo & 1u) != 0u);
    const unsigned hi = probe ? cube_plane_probe(L, tab, Xc, c.z, side) : cube_plane_scan(A, Xc, c.z, side);
    return lo | (hi << 8);
}

// ---- bricks: a second, finer level under the cubes the coarse grid cannot settle --------------------
// The uncertified cubes hug the decision surfaces (shells a few cubes thick).  Such a cube can carry a
// BRICK of 4 x 4 x 4 fine cubes (a quarter of its side), each certified on its own by the very same
// functions: the shell that stays uncertified is four times thinner.  The coarse texel (32 bit) is
// then not a result but a pointer: bit 31, the brick's number, the parity of the coarse cube's three
// indices and the cube's reach bits (bits 5-6, where the reach-only sweep reads them in either kind
// of texel).  The sweep computes the fine cube of a point with its own arithmetic; the coarse fetch
// resolves cube coordinates to 1 / 256 of a cube, so next to a cube face the texture unit may have
// picked the neighbour — the parity bits tell, and the point then belongs to the outermost fine cube
// on that side.  Every fine cube is certified over its box widened by the COARSE cube's pad (pad_h),
// which covers exactly that slack.  classic texel: bit 31 clear, low 16 bits as described above.
constexpr unsigned kVolBrick = 0x80000000u;
constexpr unsigned kBrickMaxCount = 1u << 26;
constexpr int kBrickSub = 4;  // fine cubes per axis
LRM_HD unsigned brick_texel(unsigned idx, int ix, int iy, int iz, unsigned reach_bits) {
    return kVolBrick | ((unsigned)(ix & 1) << 28) | ((unsigned)(iy & 1) << 29) | ((unsigned)(iz & 1) << 30) |
           (reach_bits & (kVolReachKnown | kVolReachValue)) | (idx & 31u) | ((idx >> 5) << 7);
}
LRM_HD unsigned brick_index(unsigned w) { return (w & 31u) | ((w >> 2) & 0x03ffffe0u); }
// does a cube with this classic texel want a brick?  (no certified choice, or no plane label)
LRM_HD bool brick_candidate(unsigned word) { return !(word & kVolPure) || (word >> 8) == 0u; }
// slot (0 .. 63, x fastest) of the fine cube that holds p in the brick the coarse fetch returned
LRM_HD unsigned brick_slot(const VolumeView& V, unsigned w, const CoxaPoint p) {
    const float inv_f = 4.f * V.inv_cell;
    const int fx = (int)floorf(fmaf(p.x, inv_f, 4.f * V.o)), fy = (int)floorf(fmaf(p.y, inv_f, 4.f * V.oy)),
              fz = (int)floorf(fmaf(p.z, inv_f, 4.f * V.o));
    // same coarse cube as the texture unit's (parity agrees): the fine index; the neighbour: the
    // point sits within the pad of the shared face
    const unsigned dx = ((unsigned)(fx >> 2) ^ (w >> 28)) & 1u, dy = ((unsigned)(fy >> 2) ^ (w >> 29)) & 1u,
                   dz = ((unsigned)(fz >> 2) ^ (w >> 30)) & 1u;
    const unsigned sx = (unsigned)fx & 3u, sy = (unsigned)fy & 3u, sz = (unsigned)fz & 3u;
    const unsigned lx = dx ? (sx >= 2u ? 0u : 3u) : sx, ly = dy ? (sy >= 2u ? 0u : 3u) : sy,
                   lz = dz ? (sz >= 2u ? 0u : 3u) : sz;
    return (lz << 4) | (ly << 2) | lx;
}
// Classic 16-bit texel of fine cube `slot` of the brick under the coarse cube at (x0, y0, z0), side h,
// whose own classic texel is `coarse` — serially (host emulation; the device build shares the work).
// A coarse cube whose choice is certified hands its low byte to its fine cubes (their boxes lie inside
// its own): only the plane label is looked for again.
LRM_HD unsigned brick_fine_word(const LegPlan& L, const SectorTable& tab, const FastTables& FT, const AtlasView& A,
                                unsigned coarse, float x0, float y0, float z0, float h, unsigned slot) {
    const float hf = h * (1.f / kBrickSub);
    const float xf = x0 + (float)(slot & 3u) * hf, yf = y0 + (float)((slot >> 2) & 3u) * hf,
                zf = z0 + (float)(slot >> 4) * hf;
    if (coarse & kVolPure) {
        float side;
        const CoxaPoint c = cube_centre(xf, yf, zf, hf, &side, h);
        const unsigned hi = cube_plane_scan(A, solution_plane_x(L, c, (coarse & 1u) != 0u), c.z, side);
        return (coarse & 0xffu) | (hi << 8);
    }
    return choice_cell_word(L, tab, FT, A, xf, yf, zf, hf, false, h);
}

// One point of a certified cube: the chosen solution only.  Same arithmetic as dist_fast for that
// solution.  Straight-line; returns 0 = done, 1 = cube uncertified (dist_fast can still decide the
// point), 2 = the chosen solution's plane cell is uncertified (full evaluation).
// dist_choice in two halves, for the streaming kernel: the front half ends with the atlas fetch,
// the back half finishes the point.  RULE = false drops the limit-plane rule (one_leg.cu:258-274),
// which can only fire when the plane point is valid (bit 6 of the cell): the kernel takes that
// version when no point of a warp's trip has a valid plane point — most of the far field — and
// gets the same bits for less work.
struct ChoiceFront {
    float cs, ss, X;
    unsigned la;
};
template <bool TEX>
LRM_HD ChoiceFront dist_choice_front(const LegPlan& L, const YawSol* sols, unsigned cube, const AtlasView& A,
                                     const CoxaPoint p) {
    const YawSol& s = sols[cube & 31u];
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float inv_rho = fast_rsqrt(rho2 > 1.0e-12f ? rho2 : 1.f);
    const float ux = p.x * inv_rho, uy = p.y * inv_rho;
    ChoiceFront f;
    f.cs = fmaf(s.k, ux, s.c_cs), f.ss = fmaf(s.k, uy, s.c_ss);
    f.X = fmaf(p.x, f.cs, p.y * f.ss) - L.coxa_length;
    f.la = atlas_fetch<TEX>(A, fmaf(f.X, A.inv_cell, A.ox), fmaf(p.z, A.inv_cell, A.oy));
    return f;
}
template <bool RULE>
LRM_HD int dist_choice_back(const LegPlan& L, const YawSol* sols, unsigned cube, const ChoiceFront& f,
                            const WinnerTable& W, const CoxaPoint p, DistResult* out) {
    const YawSol& s = sols[cube & 31u];
    const float yr = fmaf(p.y, f.cs, -p.x * f.ss);
    const PlaneResult pl = plane_from_label(W, f.la, f.X, p.z);
    float vx, vy, vz;
    bool res;
    if (RULE) {
        const float yl = fmaf(p.y, s.cl, -p.x * s.sl);
        const BranchResult b = fast_branch(p, s, f.cs, f.ss, yr, yl, pl);
        vx = b.vx, vy = b.vy, vz = b.vz, res = b.res;
    } else {
        // pl.valid is false: to_plane and res are false, the vector is the in-plane one rotated back
        vx = fmaf(pl.dx, f.cs, -yr * f.ss), vy = fmaf(pl.dx, f.ss, yr * f.cs), vz = pl.dy, res = false;
    }
    out->flag = res;
    out->reach = res & (((cube & 1u) != 0u) == (f2i(p.x) < 0));
    out->dx = fmaf(L.Mo[0], vx, fmaf(L.Mo[1], vy, L.Mo[2] * vz));
    out->dy = fmaf(L.Mo[3], vx, fmaf(L.Mo[4], vy, L.Mo[5] * vz));
    out->dz = fmaf(L.Mo[6], vx, fmaf(L.Mo[7], vy, L.Mo[8] * vz));
    return (cube & kVolPure) ? ((f.la & kAtlasPure) ? 0 : 2) : 1;
}
// Tier 0: one point whose cube texel carries both the chosen solution and its plane label.  The
// very operations of dist_choice_back on the label, without any plane-atlas fetch.  Returns 0 =
// done, 1 = cube uncertified (-> dist_fast), 3 = solution certified but the cube has no plane
// label (-> dist_choice through the plane atlas).  `word` is the 16-bit texel.
template <bool RULE>
LRM_HD int dist_choice_label(const LegPlan& L, const YawSol* sols, unsigned word, const WinnerTable& W,
                             const CoxaPoint p, DistResult* out) {
    const unsigned cube = word & 0xffu, la = word >> 8;
    const YawSol& s = sols[cube & 31u];
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    // no guard against rho2 = 0: certified cubes stay clear of the coxa axis (where the guard of
    // dist_choice is the identity), and an uncertified point's result is discarded
    const float inv_rho = fast_rsqrt(rho2);
    const float ux = p.x * inv_rho, uy = p.y * inv_rho;
    ChoiceFront f;
    f.cs = fmaf(s.k, ux, s.c_cs), f.ss = fmaf(s.k, uy, s.c_ss);
    f.X = fmaf(p.x, f.cs, p.y * f.ss) - L.coxa_length;
    f.la = la;
    dist_choice_back<RULE>(L, sols, cube, f, W, p, out);
    return (cube & kVolPure) ? ((la & kAtlasPure) ? 0 : 3) : 1;
}

template <bool TEX>
LRM_HD int dist_choice(const LegPlan& L, const YawSol* sols, unsigned cube, const AtlasView& A,
                       const WinnerTable& W, const CoxaPoint p, DistResult* out) {
    const YawSol& s = sols[cube & 31u];
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float inv_rho = fast_rsqrt(rho2 > 1.0e-12f ? rho2 : 1.f);  // certified cubes stay clear of the axis
    const float ux = p.x * inv_rho, uy = p.y * inv_rho;
    const float cs = fmaf(s.k, ux, s.c_cs), ss = fmaf(s.k, uy, s.c_ss);
    const float X = fmaf(p.x, cs, p.y * ss) - L.coxa_length;
    const unsigned la = atlas_fetch<TEX>(A, fmaf(X, A.inv_cell, A.ox), fmaf(p.z, A.inv_cell, A.oy));
    const float yr = fmaf(p.y, cs, -p.x * ss), yl = fmaf(p.y, s.cl, -p.x * s.sl);
    const BranchResult b = fast_branch(p, s, cs, ss, yr, yl, plane_from_label(W, la, X, p.z));
    // no pair has two unsaturated solutions (FastTables::both_unsat): res of the other one is
    // false wherever this one is chosen, so res || res_other = res, and reachability_circles'
    // flag is res of the solution on the point's own side of the coxa axis
    out->flag = b.res;
    out->reach = b.res & (((cube & 1u) != 0u) == (f2i(p.x) < 0));
    out->dx = fmaf(L.Mo[0], b.vx, fmaf(L.Mo[1], b.vy, L.Mo[2] * b.vz));
    out->dy = fmaf(L.Mo[3], b.vx, fmaf(L.Mo[4], b.vy, L.Mo[5] * b.vz));
    out->dz = fmaf(L.Mo[6], b.vx, fmaf(L.Mo[7], b.vy, L.Mo[8] * b.vz));
    return (cube & kVolPure) ? ((la & kAtlasPure) ? 0 : 2) : 1;
}

// The chosen solution of a certified cube with the EXPLICIT plane evaluation (plane_clamp) instead
// of the plane atlas: for points whose plane cell is uncertified.  Cannot fail.
LRM_HD void dist_choice_clamp(const LegPlan& L, const SectorTable& tab, const YawSol* sols, unsigned cube,
                              const CoxaPoint p, DistResult* out) {
    const YawSol& s = sols[cube & 31u];
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float inv_rho = fast_rsqrt(rho2 > 1.0e-12f ? rho2 : 1.f);
    const float ux = p.x * inv_rho, uy = p.y * inv_rho;
    const float cs = fmaf(s.k, ux, s.c_cs), ss = fmaf(s.k, uy, s.c_ss);
    const float X = fmaf(p.x, cs, p.y * ss) - L.coxa_length;
    const float yr = fmaf(p.y, cs, -p.x * ss), yl = fmaf(p.y, s.cl, -p.x * s.sl);
    const BranchResult b = fast_branch(p, s, cs, ss, yr, yl, plane_clamp<false>(L, tab, X, p.z));
    out->flag = b.res;
    out->reach = b.res & (((cube & 1u) != 0u) == (f2i(p.x) < 0));
    out->dx = fmaf(L.Mo[0], b.vx, fmaf(L.Mo[1], b.vy, L.Mo[2] * b.vz));
    out->dy = fmaf(L.Mo[3], b.vx, fmaf(L.Mo[4], b.vy, L.Mo[5] * b.vz));
    out->dz = fmaf(L.Mo[6], b.vx, fmaf(L.Mo[7], b.vy, L.Mo[8] * b.vz));
}

// ---- "inside the body cylinder under EVERY orientation" (pose search, positionability.cu) ---------
// The body cylinder of eliminateFarAndColliding (several_leg.cu:504-559) is fixed in the ORIENTATION
// frame: radius r around the body's z axis, heights in (z_lo, z_hi).  Its axis in the world frame is the
// third row of the orientation matrix.  If the axes of all orientations lie within an angle theta of
// the unit vector (ax, ay, az), then for a map point at offset d from the body — axial part
// a = d . axis, radial part rho, phi the angle between d and the axis — the angle between d and ANY of
// the axes lies in [phi - theta, phi + theta], the cylinder height of d is |d| cos of that angle and
// its cylinder radius |d| sin of it: extremes at the ends of the interval, or at 0 / pi (height) and
// pi / 2 (radius) where the interval contains them.  A point that keeps all three inside limits
// collides whatever the orientation.  (cm, sm) = (cos, sin) of theta; lo / hi / rad = the limits, a
// little inside; ok = 0: the axes spread too far, nothing is claimed; gate: see positionability.cu.
struct AxisCone {
    float ax, ay, az, cm, sm;
    float lo, hi, rad;
    int ok, gate;
};
LRM_HD bool cone_collides_always(const AxisCone& C, float dx, float dy, float dz) {
    const float a = fmaf(C.ax, dx, fmaf(C.ay, dy, C.az * dz));
    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    const float dn = sqrtf(d2), rho = sqrtf(fmaxf(fmaf(-a, a, d2), 0.f));
    const float lo = fmaf(rho, C.cm, a * C.sm) <= 0.f ? -dn : fmaf(a, C.cm, -rho * C.sm);
    const float hi = fmaf(rho, C.cm, -a * C.sm) <= 0.f ? dn : fmaf(a, C.cm, rho * C.sm);
    const float rad = (lo <= 0.f && hi >= 0.f) ? dn : fmaf(rho, C.cm, fabsf(a) * C.sm);
    return lo > C.lo && hi < C.hi && rad < C.rad;
}
// Host side: the cone of n cylinder axes (rows of 3 floats, not necessarily unit), limits 1.5 mm inside
// (z_lo, z_hi, radius) for the rounding of the rotations, 2 mrad on the angle.
inline void make_axis_cone(const float* axes, int n, float z_lo, float z_hi, float radius, AxisCone* C) {
    C->ok = 0, C->gate = 1, C->ax = 0.f, C->ay = 0.f, C->az = 1.f, C->cm = 1.f, C->sm = 0.f;
    C->lo = z_lo + 1.5f, C->hi = z_hi - 1.5f, C->rad = radius - 1.5f;
    double m[3] = {0, 0, 0};
    for (int o = 0; o < n; o++) m[0] += axes[3 * o], m[1] += axes[3 * o + 1], m[2] += axes[3 * o + 2];
    const double mn = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    if (!(mn > 1.0e-6 * n) || !(C->rad > 0.f)) return;
    double cmin = 1.0;
    for (int o = 0; o < n; o++) {
        const double x = axes[3 * o], y = axes[3 * o + 1], z = axes[3 * o + 2];
        cmin = fmin(cmin, (x * m[0] + y * m[1] + z * m[2]) / (mn * sqrt(x * x + y * y + z * z)));
    }
    const double theta = acos(fmax(-1.0, fmin(1.0, cmin))) + 2.0e-3;
    if (!(theta < 1.2)) return;  // beyond ~70 degrees the region is hardly more than the ball
    C->ok = 1;
    C->gate = theta > 0.05 ? 1 : 0;
    C->ax = (float)(m[0] / mn), C->ay = (float)(m[1] / mn), C->az = (float)(m[2] / mn);
    C->cm = (float)cos(theta), C->sm = (float)sin(theta);
}

// The same question for the full plan of the distance entry points (no gravity side, both coxa
// solutions count: distance_circles' flag is res || resflip): can any point within `rc` of the
// coxa-frame point p be reachable?  Conservative; `wedge` says the yaw limits span less than pi.
// Used by the body-space octree: a foothold farther than the child's half diagonal from a leg's
// workspace can neither be reached by that leg nor have its distance vector land inside the child.
LRM_HD bool leg_ball_possible(const LegPlan& L, const CoxaPoint p, float rc, bool wedge) {
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float rho = sqrtf(rho2);
    const float lo = fmaxf(L.inner.r - rc - 0.01f, 0.f), hi = L.outer.r + rc + 0.01f;
    const float xa = rho - L.coxa_length, xb = -rho - L.coxa_length;
    const float da2 = fmaf(xa, xa, p.z * p.z), db2 = fmaf(xb, xb, p.z * p.z);
    bool ring_a = da2 >= lo * lo && da2 <= hi * hi, ring_b = db2 >= lo * lo && db2 <= hi * hi;
    if (wedge && rho > rc) {
        // inner side of both limit lines through the coxa axis, see reach_ball_possible
        const float d_min = fmaf(L.cos_min, p.y, -L.sin_min * p.x), d_max = fmaf(L.sin_max, p.x, -L.cos_max * p.y);
        const float slack = rc + 0.01f;
        ring_a = ring_a && d_min >= -slack && d_max >= -slack;
        ring_b = ring_b && d_min <= slack && d_max <= slack;
    }
    return ring_a || ring_b;
}

}  // namespace lrm
