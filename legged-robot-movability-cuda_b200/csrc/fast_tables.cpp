// fast_tables.cpp — host-side construction of the yaw-sector table (FastTables, leg_plan.h).
//
// The yaw decisions of finish_finding_closest (one_leg.cu:222-234) are evaluated here with the
// very device functions the full path uses (yaw_tests_both, leg_math.cuh), on sample directions of
// every diamond-angle bin; a bin is certified only if all samples of the padded bin agree.
#include <cmath>
#include <cstring>
#include <limits>

#include "leg_math.cuh"
#include "leg_plan.h"

namespace lrm {

namespace {

// direction with diamond angle d in [-2, 2]
void diamond_dir(double d, double* x, double* y) {
    if (d > 1.0) {
        *y = 2.0 - d, *x = -(1.0 - *y);
    } else if (d < -1.0) {
        *y = -2.0 - d, *x = -(1.0 + *y);
    } else {
        *y = d, *x = 1.0 - std::fabs(d);
    }
}

// solution index: 0-6 direct, 7-13 flipped: [free lo, free hi, min lo, min hi, max lo, max hi, mega]
int sol_index(const YawFlags& f, bool flipped) {
    int k;
    if (f.mega) k = 6;
    else if (f.under) k = 2 + (f.upper_lim ? 1 : 0);
    else if (f.over) k = 4 + (f.upper_lim ? 1 : 0);
    else k = f.upper_lim ? 1 : 0;
    return k + (flipped ? 7 : 0);
}

int code_of(const LegPlan& L, float x, float y) {
    YawFlags fa, fb;
    yaw_tests_both(L, x, y, fa, fb);
    // same rule as dist_coxa_frame: a mega-saturated solution next to an unsaturated one is skipped
    const bool skip_a = fa.mega & !(fb.mega | fb.over | fb.under);
    const bool skip_b = fb.mega & !(fa.mega | fa.over | fa.under);
    const int ia = skip_a ? kYawSkip : sol_index(fa, false);
    const int ib = skip_b ? kYawSkip : sol_index(fb, true);
    return (ib << 4) | ia;
}

}  // namespace

void build_fast_tables(const LegPlan& L, FastTables* out) {
    std::memset(out, 0, sizeof *out);
    const float inf = std::numeric_limits<float>::infinity();
    for (int flipped = 0; flipped < 2; flipped++) {
        const float sigma = flipped ? -1.f : 1.f;
        for (int k = 0; k < 7; k++) {
            YawSol& s = out->sol[k + 7 * flipped];
            const bool hi = (k & 1) != 0;  // upper_lim
            s.cl = hi ? L.cos_max : L.cos_min;
            s.sl = hi ? L.sin_max : L.sin_min;
            s.big = 0.f, s.nsat = 0.f, s.k = 0.f, s.c_cs = 0.f, s.c_ss = 0.f, s.pad = 0.f;
            if (k < 2) {  // unsaturated: the plane of the point itself
                s.k = sigma, s.nsat = 1.f;
            } else if (k < 4) {
                s.c_cs = L.cos_min, s.c_ss = L.sin_min;
            } else if (k < 6) {
                s.c_cs = L.cos_max, s.c_ss = L.sin_max;
            } else {  // beyond limit +- pi/2: yaw -+ pi, limit-plane rule off
                s.k = -sigma, s.big = inf;
            }
        }
    }
    // pad: a quarter bin on each side (float error of the device's bin index is ~1e-6 bins)
    const double width = 4.0 / kYawBins, pad = 0.25 * width;
    const int samples = 48;
    for (int b = 0; b <= kYawBins; b++) {
        const double lo = -2.0 + b * width - pad, hi = -2.0 + (b + 1) * width + pad;
        // the seam (d = +-2) and the positive x axis (d = 0) carry the signed-zero rules of
        // atan2f: never certified
        bool ok = lo > -2.0 && hi < 2.0 && !(lo <= 0.0 && hi >= 0.0);
        int code = -1;
        for (int s = 0; ok && s <= samples; s++) {
            double x, y;
            diamond_dir(lo + (hi - lo) * s / samples, &x, &y);
            for (int scale = 0; ok && scale < 2; scale++) {  // the tests are scale-free; check two radii
                const float r = scale ? 700.f : 3.f;
                const int c = code_of(L, (float)(x * r), (float)(y * r));
                if (code < 0) code = c;
                ok = ok && c == code;
            }
        }
        out->code[b] = ok ? (uint8_t)code : kYawImpure;
    }
}

}  // namespace lrm
