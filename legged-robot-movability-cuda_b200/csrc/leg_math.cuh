// leg_math.cuh — per-point device math of the one-leg reachability / distance path.
//
// Same decisions and the same nearest-boundary construction as the reference
// (reachability_circles one_leg.cu:280-319, distance_circles :321-341, finish_finding_closest
// :215-278, eval_plane_circles :167-208, multi_circle_clamp :91-145, find_region
// circles.cu.h:48-78), re-derived so that nothing leg-constant is evaluated per point:
//   * no atan2f / sincosf: every angle comparison is a cross-product sign test against a constant
//     direction (AngleTest), and the coxa rotation uses the normalised (x, y) itself;
//   * the 4 circles of a sector come from a 4-sector table (shared memory, float4 per circle)
//     instead of being rebuilt from 8+ sin/cos per point into local memory;
//   * circle validity is a compare on squared distances with the +-CIRCLE_MARGIN folded into the
//     threshold; corner points are constants.
// Everything lives in registers; the only memory traffic is the point itself.
#pragma once
#include <cuda_runtime.h>

#include "leg_plan.h"

namespace lrm {

constexpr float kMarginF = 0.001f;  // CIRCLE_MARGIN, settings.h:9

// Sector table staged in shared memory by every CTA: [sector = upper*2 + ext][slot 1..3].
struct SectorTable {
    float4 circle[4][3];  // cx, cy, r, sgn
    float thr_s[4][4];    // sgn * (r + sgn*eps)^2 ; 4th column pads the row to 16 B
};

__device__ __forceinline__ void fill_sector_table(const LegPlan& L, SectorTable* tab, int tid,
                                                  int nthreads) {
    for (int i = tid; i < 12; i += nthreads) {
        const int sector = i / 3, j = i % 3;
        const int upper = sector >> 1, ext = sector & 1;
        PlanCircle c = L.slot[upper][j];
        if (ext && L.att_slot[upper] == j) c = L.outer;
        tab->circle[sector][j] = make_float4(c.cx, c.cy, c.r, c.sgn);
        tab->thr_s[sector][j] = c.thr_s;
    }
}

__device__ __forceinline__ bool angle_gt(const AngleTest& t, float X, float Y) {
    const float cr = fmaf(t.c, Y, fmaf(t.ns, X, t.bias));
    const bool up = (__float_as_int(Y) >= 0);  // !signbit(Y)
    const bool pos = cr > 0.f;
    const bool lower = t.lower != 0;  // uniform; bitwise forms keep this branch-free
    return (up & pos) | (lower & (up | pos));
}

__device__ __forceinline__ int find_sector(const LegPlan& L, float X, float Y) {
    const bool upper = angle_gt(L.middle, X, Y);
    const bool more = upper ? angle_gt(L.sat[1], X, Y) : angle_gt(L.sat[0], X, Y);
    const bool ext = upper != more;  // circles.cu.h:73-74
    return (upper ? 2 : 0) | (ext ? 1 : 0);
}

// one_leg.cu:65-89 on squared distances: sgn*|P-c|^2 < thr_s
__device__ __forceinline__ bool circle_ok(float cx, float cy, float sgn, float thr_s, float x,
                                          float y) {
    const float vx = x - cx, vy = y - cy;
    return sgn * fmaf(vx, vx, vy * vy) < thr_s;
}

// eval_plane_circles<REACH_USECASE>, (X, Y) already relative to the femur joint.
__device__ __forceinline__ bool plane_reach(const LegPlan& L, const SectorTable& tab, float X,
                                            float Y) {
    const int s = find_sector(L, X, Y);
    bool ok = L.inner.sgn * fmaf(X, X, Y * Y) < L.inner.thr_s;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 c = tab.circle[s][j];
        ok = ok && circle_ok(c.x, c.y, c.w, tab.thr_s[s][j], X, Y);
    }
    return ok;
}

struct PlaneResult {
    bool valid;    // the query point satisfies all 4 circles
    float dx, dy;  // P - nearest valid boundary candidate
};

// eval_plane_circles<DIST_USECASE> = insert_circles + insert_intersecv2 + multi_circle_clamp.
__device__ __forceinline__ PlaneResult plane_clamp(const LegPlan& L, const SectorTable& tab,
                                                   float X, float Y) {
    const int s = find_sector(L, X, Y);
    float cx[4], cy[4], r[4], sg[4], th[4];
    cx[0] = 0.f, cy[0] = 0.f, r[0] = L.inner.r, sg[0] = L.inner.sgn, th[0] = L.inner.thr_s;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 c = tab.circle[s][j];
        cx[j + 1] = c.x, cy[j + 1] = c.y, r[j + 1] = c.z, sg[j + 1] = c.w;
        th[j + 1] = tab.thr_s[s][j];
    }

    // project P on each circle (force_clamp_on_circle, one_leg.cu:42-63)
    float px[4], py[4], d[4];
    bool valid = true;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float vx = X - cx[j], vy = Y - cy[j];
        const float m2 = fmaf(vx, vx, vy * vy);
        float rinv = rsqrtf(m2);
        float m = m2 * rinv;
        if (!(m >= kMarginF)) {  // also catches m2 == 0 (rinv = inf, m = NaN)
            m = m2 > 0.f ? m : 0.f;
            vx = 1.f, vy = 0.f, rinv = 1.f;
        }
        d[j] = r[j] - m;
        valid = valid && (sg[j] * d[j] > -kMarginF);  // (d >= 0) == attractive, or |d| < margin
        const float k = r[j] * rinv;
        px[j] = fmaf(vx, k, cx[j]);
        py[j] = fmaf(vy, k, cy[j]);
    }

    // a projection only counts if it satisfies the other circles (one_leg.cu:122-123); the
    // first strictly closer one wins (:133-140)
    float best_abs = 999999999999999.9f, bx = 0.f, by = 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k != j) ok = ok && circle_ok(cx[k], cy[k], sg[k], th[k], px[j], py[j]);
        const float a = fabsf(d[j]);
        if (ok && best_abs > a) best_abs = a, bx = px[j], by = py[j];
    }
    // corner points compete only when P itself is outside (one_leg.cu:109-118)
    if (!valid) {
        float best2 = best_abs * best_abs;
#pragma unroll
        for (int i = 0; i < kMaxCorners; i++) {
            if (i < L.n_corners) {
                const float wx = X - L.corner_x[i], wy = Y - L.corner_y[i];
                const float w2 = fmaf(wx, wx, wy * wy);
                if (best2 > w2) best2 = w2, bx = L.corner_x[i], by = L.corner_y[i];
            }
        }
    }
    PlaneResult out;
    out.valid = valid;
    out.dx = X - bx;
    out.dy = Y - by;
    return out;
}

struct CoxaPoint {
    float x, y, z;  // point in the coxa frame
};

__device__ __forceinline__ CoxaPoint to_coxa_frame(const LegPlan& L, float x, float y, float z) {
    CoxaPoint p;
    p.x = fmaf(L.M[0], x, fmaf(L.M[1], y, fmaf(L.M[2], z, L.t[0])));
    p.y = fmaf(L.M[3], x, fmaf(L.M[4], y, fmaf(L.M[5], z, L.t[1])));
    p.z = fmaf(L.M[6], x, fmaf(L.M[7], y, fmaf(L.M[8], z, L.t[2])));
    return p;
}

// reachability_circles, one_leg.cu:280-319
__device__ __forceinline__ bool reach_coxa_frame(const LegPlan& L, const SectorTable& tab,
                                                 const CoxaPoint p) {
    const bool flip = __float_as_int(p.x) < 0;  // signbit: mirrored through the coxa axis
    const float xf = flip ? -p.x : p.x;
    const float yf = flip ? -p.y : p.y;
    if (angle_gt(L.over, xf, yf) || angle_gt(L.under, xf, -yf)) return false;
    const float rho = sqrtf(fmaf(p.x, p.x, p.y * p.y));
    const float X = (flip ? -rho : rho) - L.coxa_length;
    return plane_reach(L, tab, X, p.z);
}

struct BranchResult {
    bool res;          // was_valid && !coxa_saturated
    float vx, vy, vz;  // vector in the coxa frame
    float n2;          // its squared norm
};

// finish_finding_closest<bool>, one_leg.cu:215-278, for the coxa solution whose yaw is the angle
// of (wx, wy): (wx, wy) = (x, y) for the direct solution, (-x, 0 - y) for the flipped one.
__device__ __forceinline__ BranchResult closest_for_branch(const LegPlan& L, const SectorTable& tab,
                                                           const CoxaPoint p, float wx, float wy,
                                                           float inv_rho) {
    const bool mega = angle_gt(L.mega_hi, wx, wy) || angle_gt(L.mega_lo, wx, -wy);
    const bool over = angle_gt(L.over, wx, wy);
    const bool under = angle_gt(L.under, wx, -wy);
    // unit direction of the saturated yaw
    float cs = wx * inv_rho, ss = wy * inv_rho;
    if (inv_rho == 0.f) cs = 1.f, ss = 0.f;  // point on the coxa axis: yaw 0
    if (mega) {
        cs = -cs, ss = -ss;  // yaw -+ pi
    } else if (under) {
        cs = L.cos_min, ss = L.sin_min;
    } else if (over) {
        cs = L.cos_max, ss = L.sin_max;
    }
    const bool saturated = mega || over || under;
    const float xr = fmaf(p.x, cs, p.y * ss);
    const float yr = fmaf(p.y, cs, -p.x * ss);

    const PlaneResult pl = plane_clamp(L, tab, xr - L.coxa_length, p.z);
    float ux = pl.dx, uy = yr, uz = pl.dy;  // in the saturated-yaw frame
    float n2 = fmaf(ux, ux, fmaf(uy, uy, uz * uz));

    BranchResult out;
    out.res = pl.valid && !saturated;
    // in-plane region reached but a coxa-limit half-plane is nearer (one_leg.cu:258-274)
    const bool upper_lim = angle_gt(L.mid, wx, wy);
    const float cl = upper_lim ? L.cos_max : L.cos_min;
    const float sl = upper_lim ? L.sin_max : L.sin_min;
    const float yl = fmaf(p.y, cl, -p.x * sl);
    if (pl.valid && !mega && n2 > yl * yl) {
        out.vx = -yl * sl, out.vy = yl * cl, out.vz = 0.f, out.n2 = yl * yl;
    } else {
        out.vx = fmaf(ux, cs, -uy * ss);
        out.vy = fmaf(ux, ss, uy * cs);
        out.vz = uz;
        out.n2 = n2;
    }
    return out;
}

struct DistResult {
    bool flag;        // distance_circles' return: res || resflip
    bool reach;       // reachability_circles of the same point
    float dx, dy, dz; // world-frame vector
};

// distance_circles (one_leg.cu:321-341) + the way back to the world frame.
__device__ __forceinline__ DistResult dist_coxa_frame(const LegPlan& L, const SectorTable& tab,
                                                      const CoxaPoint p) {
    const float rho2 = fmaf(p.x, p.x, p.y * p.y);
    const float inv_rho = rho2 > 0.f ? rsqrtf(rho2) : 0.f;
    const BranchResult a = closest_for_branch(L, tab, p, p.x, p.y, inv_rho);
    // flipped yaw = yaw -+ pi: the angle of (-x, -y); "0 - y" keeps atan2f's +pi (not -pi) for
    // y = +0, x > 0, like coxangle + pi does in the reference (one_leg.cu:329)
    const BranchResult b = closest_for_branch(L, tab, p, -p.x, 0.f - p.y, inv_rho);
    const bool direct = (a.res == b.res) ? (a.n2 < b.n2) : a.res;
    const float vx = direct ? a.vx : b.vx, vy = direct ? a.vy : b.vy, vz = direct ? a.vz : b.vz;
    DistResult out;
    out.flag = a.res || b.res;
    out.reach = (__float_as_int(p.x) < 0) ? b.res : a.res;
    out.dx = fmaf(L.Mo[0], vx, fmaf(L.Mo[1], vy, L.Mo[2] * vz));
    out.dy = fmaf(L.Mo[3], vx, fmaf(L.Mo[4], vy, L.Mo[5] * vz));
    out.dz = fmaf(L.Mo[6], vx, fmaf(L.Mo[7], vy, L.Mo[8] * vz));
    return out;
}

}  // namespace lrm
