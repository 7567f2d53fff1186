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

YawSol make_sol(const LegPlan& L, int kind, bool flipped) {
    const float inf = std::numeric_limits<float>::infinity();
    const float sigma = flipped ? -1.f : 1.f;
    YawSol s;
    std::memset(&s, 0, sizeof s);
    if (kind == kYawSkipped) {
        // never evaluated; harmless finite values
        s.k = sigma, s.cl = L.cos_min, s.sl = L.sin_min, s.big = inf, s.present = 0.f;
        return s;
    }
    const bool hi = (kind & 1) != 0 && kind < 6;
    s.cl = hi ? L.cos_max : L.cos_min;
    s.sl = hi ? L.sin_max : L.sin_min;
    s.present = 1.f;
    if (kind < 2) {  // unsaturated: the plane of the point itself
        s.k = sigma, s.nsat = 1.f;
    } else if (kind < 4) {
        s.c_cs = L.cos_min, s.c_ss = L.sin_min;
    } else if (kind < 6) {
        s.c_cs = L.cos_max, s.c_ss = L.sin_max;
    } else {  // beyond limit +- pi/2: yaw -+ pi, limit-plane rule off
        s.k = -sigma, s.big = inf;
    }
    return s;
}

}  // namespace

void build_fast_tables(const LegPlan& L, FastTables* out) {
    std::memset(out, 0, sizeof *out);
    int combos[kYawPairs];
    int ncombo = 0;
    // pad: a quarter bin on each side (float error of the device's bin index is ~1e-6 bins)
    const double width = 4.0 / kYawBins, pad = 0.25 * width;
    const int samples = 48;
    for (int b = 0; b <= kYawBins; b++) {
        const double lo = -2.0 + b * width - pad, hi = -2.0 + (b + 1) * width + pad;
        // the seam (d = +-2) and the positive x axis (d = 0) carry the signed-zero rules of
        // atan2f: never certified
        bool ok = lo > -2.0 && hi < 2.0 && !(lo <= 0.0 && hi >= 0.0);
        int combo = -1;
        for (int s = 0; ok && s <= samples; s++) {
            double x, y;
            diamond_dir(lo + (hi - lo) * s / samples, &x, &y);
            for (int scale = 0; ok && scale < 2; scale++) {  // the tests are scale-free; check two radii
                const float r = scale ? 700.f : 3.f;
                const int c = yaw_combo(L, (float)(x * r), (float)(y * r));
                if (combo < 0) combo = c;
                ok = ok && c == combo;
            }
        }
        int id = -1;
        if (ok) {
            for (int i = 0; i < ncombo; i++)
                if (combos[i] == combo) id = i;
            if (id < 0 && ncombo < kYawPairs) {
                id = ncombo;
                combos[ncombo++] = combo;
                out->pair[id].a = make_sol(L, combo & 7, false);
                out->pair[id].b = make_sol(L, combo >> 3, true);
            }
        }
        out->code[b] = id >= 0 ? (uint8_t)id : kYawImpure;  // table full (exotic leg): full evaluation
    }
    out->ncombo = ncombo;
    for (int i = 0; i < ncombo; i++) {
        out->combo[i] = combos[i];
        if (out->pair[i].a.present != 0.f && out->pair[i].b.present != 0.f && out->pair[i].a.nsat != 0.f &&
            out->pair[i].b.nsat != 0.f)
            out->both_unsat = 1;
    }
}

}  // namespace lrm
