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

int emu_plan_is_generic(const lrm_leg_t* leg, const float* quat) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    return L.generic;
}

int emu_sizeof_plan() { return (int)sizeof(lrm::LegPlan); }

}  // extern "C"
