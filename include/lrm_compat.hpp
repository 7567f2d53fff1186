// lrm_compat.hpp — the reference's C++ plugin surface re-exported on top of the C ABI (lrm_c.h).
//
// A plain-C++ caller of the reference (bench.cpp:129-158, several_leg.cpp:143-148,183-186) includes
// cross_compiled.cuh / one_leg.cu.h and calls
//     apply_kernel(points, dim, reachability_global_kernel, out)      -> float ms
//     apply_kernel(points, dim, distance_global_kernel,     out)      -> float ms
// with Array<float3> / Array<bool> / LegDimensions.  Including THIS header instead (and linking
// liblrm_b200.so) keeps those call sites compiling unchanged: the kernel "handles" are tag objects,
// dispatch is by overload instead of by function pointer, layouts and ownership are the reference's
// (caller owns host arrays; the call stages through device memory and returns kernel-only
// milliseconds; errors print and exit(EXIT_FAILURE) like CUDA_CHECK_ERROR, cross_compiled.cu:12-20).
//
// The CPU twins apply_reach_cpu / apply_dist_cpu (cross_compiled.cuh:12-15) are deliberately NOT
// provided: the product has no CPU path.  They stay with the reference (or with oracle/ in tests).
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <tuple>

#include "lrm_c.h"

#ifndef __VECTOR_TYPES_H__  // allow use without the CUDA headers
struct float3 {
    float x, y, z;
};
#endif

typedef lrm_leg_t LegDimensions;  // HeaderCPP.h:19-52, same 14 floats

template <typename T>
struct Array {  // HeaderCUDA.h:38-66
    size_t length;
    T* elements;
};

namespace lrm_compat {
struct ReachKernelTag {};
struct DistKernelTag {};
struct ForwardKineTag {};
inline void die(const char* where, int rc) {
    std::fprintf(stderr, "CUDA error in %s: %s (lrm status %d)\n", where, lrm_last_error(), rc);
    std::exit(EXIT_FAILURE);
}
}  // namespace lrm_compat

// kernel handles, one_leg.cu.h:18-40
static const lrm_compat::ReachKernelTag reachability_global_kernel{}, reachability_circles_kernel{};
static const lrm_compat::DistKernelTag distance_global_kernel{}, distance_circles_kernel{};
static const lrm_compat::ForwardKineTag forward_kine_kernel{};

// cross_compiled.cuh:4-7
inline float apply_kernel(const Array<float3> input, const LegDimensions dim,
                          lrm_compat::ReachKernelTag, Array<bool> const output) {
    static_assert(sizeof(bool) == 1, "Array<bool> is one byte per flag");
    float ms = 0.f;
    int rc = lrm_reach(&input.elements->x, input.length, &dim, nullptr,
                       reinterpret_cast<uint8_t*>(output.elements), 0, nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("reachability kernel", rc);
    return ms;
}
inline float apply_kernel(const Array<float3> input, const LegDimensions dim,
                          lrm_compat::DistKernelTag, Array<float3> const output) {
    float ms = 0.f;
    int rc = lrm_dist(&input.elements->x, input.length, &dim, nullptr, &output.elements->x, nullptr,
                      0, nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("distance kernel", rc);
    return ms;
}
inline float apply_kernel(const Array<float3> input, const LegDimensions dim,
                          lrm_compat::ForwardKineTag, Array<float3> const output) {
    float ms = 0.f;
    int rc = lrm_forward_kine(&input.elements->x, input.length, &dim, &output.elements->x, 0,
                              nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("forward_kine kernel", rc);
    return ms;
}

// static_variables.h
inline LegDimensions get_moonbot_leg(float azimut) {
    LegDimensions l;
    lrm_default_leg(0, azimut, &l);
    return l;
}
inline LegDimensions get_M2_leg(float azimut) {
    LegDimensions l;
    lrm_default_leg(1, azimut, &l);
    return l;
}

// several_leg.cu.h:13-15 — returns new[]-allocated host arrays the caller delete[]s
// (several_leg.cpp:119-121).  Standable bodies keep their input order (the reference's order is
// whatever thrust::partition leaves); the count array is the reference's dummy 3s
// (several_leg.cu:867-868).
inline std::tuple<Array<float3>, Array<int>> robot_full_struct(Array<float3> body_map,
                                                               Array<float3> target_map,
                                                               Array<LegDimensions> legs) {
    float quats[45 * 4];
    lrm_full_struct_orientations(quats, 45);
    uint8_t* flags = new uint8_t[body_map.length ? body_map.length : 1];
    lrm_posit_opts_t opts = {1, 0};
    int rc = lrm_positionability(&body_map.elements->x, body_map.length, &target_map.elements->x,
                                 target_map.length, legs.elements, (int)legs.length, quats, 45, &opts,
                                 flags, 0, nullptr, nullptr);
    if (rc != LRM_OK) lrm_compat::die("robot_full_struct", rc);
    size_t n = 0;
    for (size_t i = 0; i < body_map.length; i++) n += flags[i] != 0;
    Array<float3> out_body{n, new float3[n ? n : 1]};
    Array<int> out_count{n, new int[n ? n : 1]};
    size_t k = 0;
    for (size_t i = 0; i < body_map.length; i++)
        if (flags[i]) {
            out_body.elements[k] = body_map.elements[i];
            out_count.elements[k] = 3;
            k++;
        }
    delete[] flags;
    return std::make_tuple(out_body, out_count);
}

// several_leg_octree.cu.h:4 — apply_oct(footholds, dim, output&): like the reference it
// delete[]s the caller's output.elements and replaces it with a new[] array of the valid leaf
// centres (several_leg_octree.cu:469-472), and returns the elapsed kernel milliseconds.
// MAX_DEPTH is a compile-time constant in the reference (settings.h:15, shipped as 1).
#ifndef LRM_COMPAT_MAX_DEPTH
#define LRM_COMPAT_MAX_DEPTH 1
#endif
inline float apply_oct(Array<float3> input, LegDimensions dim, Array<float3>& output) {
    float ms = 0.f;
    size_t count = 0, cap = 1024;
    float* buf = new float[3 * cap];
    int rc = lrm_oct(&input.elements->x, input.length, &dim, LRM_COMPAT_MAX_DEPTH, buf, cap, &count, 0,
                     nullptr, &ms);
    if (rc == LRM_OK && count > cap) {
        delete[] buf;
        cap = count;
        buf = new float[3 * cap];
        rc = lrm_oct(&input.elements->x, input.length, &dim, LRM_COMPAT_MAX_DEPTH, buf, cap, &count, 0,
                     nullptr, &ms);
    }
    if (rc != LRM_OK) lrm_compat::die("apply_oct", rc);
    delete[] output.elements;
    output.length = count;
    output.elements = new float3[count ? count : 1];
    for (size_t i = 0; i < count; i++) output.elements[i] = {buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]};
    delete[] buf;
    return ms;
}

// cross_compiled.cuh:9-10 — apply_recurs<float3, LegDimensions, float3>(input, dim, output)
template <typename T_in = float3, typename param = LegDimensions, typename T_out = float3>
inline float apply_recurs(const Array<float3> input, const LegDimensions dim, Array<float3> const output) {
    float ms = 0.f;
    int rc = lrm_recurs(&input.elements->x, input.length, &dim, nullptr, LRM_COMPAT_MAX_DEPTH,
                        &output.elements->x, 0, nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("apply_recurs", rc);
    return ms;
}
