// TEST INFRASTRUCTURE ONLY — host-side emulator of the device math.
//
// Compiles legged-robot-movability-cuda_b200/csrc/leg_math.cuh (the exact source the CUDA kernels
// are built from) for the HOST, so that the CPU test suite can compare the product's per-point
// arithmetic and its leg-plan construction against the oracle without a GPU.  It is not part of
// liblrm_b200.so, is not reachable from the C ABI, and is never used by bench.py: it exists to
// catch algorithmic regressions before GPU minutes are spent.  Differences from the device:
// rsqrtf is 1/sqrtf here (a few ulp), everything else is the same FP32 expression tree.
#include <cstdint>
#include <cstring>

#include "leg_math.cuh"
#include "leg_plan.h"

namespace {
void host_table(const lrm::LegPlan& L, lrm::SectorTable* tab) {
    lrm::fill_sector_table(L, tab, 0, 1);
}
}  // namespace

extern "C" {

void emu_reach(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, uint8_t* out) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        out[i] = lrm::reach_coxa_frame(L, tab, p) ? 1 : 0;
    }
}

void emu_dist(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, float* out_vec,
              uint8_t* out_flag, uint8_t* out_reach) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const lrm::DistResult r = L.generic ? lrm::dist_coxa_frame<true>(L, tab, p) : lrm::dist_coxa_frame<false>(L, tab, p);
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
        if (out_flag) out_flag[i] = r.flag ? 1 : 0;
        if (out_reach) out_reach[i] = r.reach ? 1 : 0;
    }
}

// the reachable_rotate_leg predicate as the positionability kernel evaluates it
void emu_leg_reaches(const float* offsets, size_t n, const lrm_leg_t* leg, const float* quat,
                     uint8_t* out) {
    lrm::LegPlan L;
    lrm::build_leg_plan_rotated_limits(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const float vx = offsets[3 * i], vy = offsets[3 * i + 1], vz = offsets[3 * i + 2];
        const float g = fmaf(L.grav[0], vx, fmaf(L.grav[1], vy, L.grav[2] * vz));
        out[i] = (!(g < 0.f) && lrm::reach_coxa_frame(L, tab, lrm::to_coxa_frame(L, vx, vy, vz))) ? 1 : 0;
    }
}

// same points through the explicit cross-validation path (the fallback for exotic legs)
void emu_dist_generic(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat,
                      float* out_vec) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const lrm::DistResult r = lrm::dist_coxa_frame<true>(L, tab, p);
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
    }
}

// The atlas path exactly as the streaming kernel runs it: certified cells take plane_lookup, the
// rest fall back to the full evaluation.  Returns the number of points that fell back.
size_t emu_dist_atlas(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, int dim,
                      float cell, float* out_vec, uint8_t* out_flag, uint8_t* out_reach,
                      size_t* pure_cells) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    lrm::WinnerTable win;
    lrm::fill_winner_table(L, &win, 0, 1);
    const float origin = -0.5f * dim * cell;
    signed char* cells = new signed char[(size_t)dim * dim];
    const float need = cell * 0.70711f * 1.02f + 2.0e-3f;  // same rule as atlas_build_kernel
    size_t pure = 0;
    for (int iy = 0; iy < dim; iy++)
        for (int ix = 0; ix < dim; ix++) {
            const float X = origin + ((float)ix + 0.5f) * cell, Y = origin + ((float)iy + 0.5f) * cell;
            const lrm::PlaneProbe pr = lrm::plane_probe(L, tab, X, Y);
            const bool ok = pr.safety > need;
            cells[lrm::atlas_index(dim, ix, iy)] = (signed char)(ok ? pr.label : 0x80);
            pure += ok;
        }
    if (pure_cells) *pure_cells = pure;
    lrm::AtlasView A{cells, origin, origin, 1.0f / cell, dim, dim};
    size_t fallback = 0;
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        bool ok = true;
        lrm::DistResult r = lrm::dist_coxa_frame<false, true>(L, tab, p, &A, &win, &ok);
        if (!ok) {
            r = lrm::dist_coxa_frame<false, false>(L, tab, p);
            fallback++;
        }
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
        if (out_flag) out_flag[i] = r.flag ? 1 : 0;
        if (out_reach) out_reach[i] = r.reach ? 1 : 0;
    }
    delete[] cells;
    return fallback;
}

int emu_plan_is_generic(const lrm_leg_t* leg, const float* quat) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    return L.generic;
}

int emu_sizeof_plan() { return (int)sizeof(lrm::LegPlan); }

}  // extern "C"
