// TEST / BENCH INFRASTRUCTURE ONLY — never linked into the product library.
//
// extern "C" window onto the reference's own GPU entry points, compiled UNMODIFIED from
// /root/reference for sm_100 (oracle/Makefile `make refgpu` -> oracle/_ref/libref_gpu.so):
//   apply_kernel<float3,LegDimensions,bool|float3>   cross_compiled.cu:34-79
//   apply_recurs                                      cross_compiled.cu:82-139
//   apply_oct                                         several_leg_octree.cu:391-488
// Used on the GPU box as (i) a second parity column next to the CPU oracle and (ii) the
// "reference kernels recompiled for sm_100" timing column of tools/bench_vs_refgpu.py.
// Nothing is copied from the reference; this file only forwards.
#include "HeaderCPP.h"
#include "HeaderCUDA.h"
#include "cross_compiled.cuh"
#include "one_leg.cu.h"
#include "several_leg_octree.cu.h"

#include <cstdint>
#include <cstring>

namespace {
inline LegDimensions leg_from(const float* l) {
    LegDimensions d;
    static_assert(sizeof(LegDimensions) == 14 * sizeof(float), "LegDimensions layout");
    std::memcpy(&d, l, sizeof(d));
    return d;
}
}  // namespace

extern "C" {

// reachability_global_kernel through apply_kernel; returns the reference's kernel-only ms
float refgpu_reach(const float* xyz, size_t n, const float* leg14, uint8_t* out) {
    static_assert(sizeof(bool) == 1, "bool layout");
    Array<float3> in{n, (float3*)xyz};
    Array<bool> o{n, (bool*)out};
    return apply_kernel(in, leg_from(leg14), reachability_global_kernel, o);
}

// distance_global_kernel through apply_kernel
float refgpu_dist(const float* xyz, size_t n, const float* leg14, float* out_xyz) {
    Array<float3> in{n, (float3*)xyz};
    Array<float3> o{n, (float3*)out_xyz};
    return apply_kernel(in, leg_from(leg14), distance_global_kernel, o);
}

// apply_recurs: paints (depth, 0, 0) on the points inside leaf boxes; out must be pre-filled
float refgpu_recurs(const float* xyz, size_t n, const float* leg14, float* out_xyz) {
    Array<float3> in{n, (float3*)xyz};
    Array<float3> o{n, (float3*)out_xyz};
    return apply_recurs<float3, LegDimensions, float3>(in, leg_from(leg14), o);
}

// apply_oct: the reference delete[]s output.elements and replaces it; copy out and free here
float refgpu_oct(const float* footholds, size_t nt, const float* leg14, float* out_xyz, size_t cap,
                 size_t* count) {
    Array<float3> in{nt, (float3*)footholds};
    Array<float3> o{1, new float3[1]};
    const float ms = apply_oct(in, leg_from(leg14), o);
    *count = o.length;
    const size_t m = o.length < cap ? o.length : cap;
    if (m) std::memcpy(out_xyz, o.elements, m * sizeof(float3));
    delete[] o.elements;
    return ms;
}

}  // extern "C"
