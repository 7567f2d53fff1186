// lrm_compat.hpp — the reference's C++ plugin surface re-exported on top of the C ABI (lrm_c.h).
//
// A plain-C++ caller of the reference (bench.cpp:129-158, several_leg.cpp:143-148,183-186) includes
// cross_compiled.cuh / one_leg.cu.h and calls
//     apply_kernel(points, dim, reachability_global_kernel, out)      -> float ms
//     apply_kernel(points, dim, distance_global_kernel,     out)      -> float ms
// with Array<float3> / Array<bool> / LegDimensions.  Including THIS header instead (and linking
// liblrm_b200.so) keeps those call sites compiling unchanged: the kernel handles are host functions
// with the reference's own signatures, apply_kernel is the reference's template (dispatching on the
// handle's address), layouts and ownership are the reference's (caller owns host arrays; the call
// stages through device memory and returns kernel-only milliseconds; errors print and
// exit(EXIT_FAILURE) like CUDA_CHECK_ERROR, cross_compiled.cu:12-20).
//
// The CPU twins apply_reach_cpu / apply_dist_cpu (cross_compiled.cuh:12-15) are declared but
// deliberately NOT defined: the product has no CPU path.
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
inline void die(const char* where, int rc) {
    std::fprintf(stderr, "CUDA error in %s: %s (lrm status %d)\n", where, lrm_last_error(), rc);
    std::exit(EXIT_FAILURE);
}
inline float run_reach(const Array<float3> input, const LegDimensions dim, Array<bool> const output) {
    static_assert(sizeof(bool) == 1, "Array<bool> is one byte per flag");
    float ms = 0.f;
    int rc = lrm_reach(&input.elements->x, input.length, &dim, nullptr,
                       reinterpret_cast<uint8_t*>(output.elements), 0, nullptr, &ms);
    if (rc != LRM_OK) die("reachability kernel", rc);
    return ms;
}
inline float run_dist(const Array<float3> input, const LegDimensions dim, Array<float3> const output) {
    float ms = 0.f;
    int rc = lrm_dist(&input.elements->x, input.length, &dim, nullptr, &output.elements->x, nullptr,
                      0, nullptr, &ms);
    if (rc != LRM_OK) die("distance kernel", rc);
    return ms;
}
inline float run_forward_kine(const Array<float3> input, const LegDimensions dim, Array<float3> const output) {
    float ms = 0.f;
    int rc = lrm_forward_kine(&input.elements->x, input.length, &dim, &output.elements->x, 0,
                              nullptr, &ms);
    if (rc != LRM_OK) die("forward_kine kernel", rc);
    return ms;
}
}  // namespace lrm_compat

// Kernel handles with the reference's own names AND signatures (one_leg.cu.h:18-40): host symbols,
// so that `apply_kernel(points, dim, reachability_global_kernel, out)`, the explicit form
// `apply_kernel<float3, LegDimensions, bool>(...)` and call sites that keep the kernel in a
// function-pointer variable all compile and link unchanged.  apply_kernel dispatches on pointer
// identity (inline functions have one address per program).  Calling a handle directly runs the
// sweep with the same host-array contract.  The *_circles_* kernels are the identity-orientation
// forms (one_leg.cu:343-375): same result as the global ones with quatTest = {1,0,0,0}.
inline void reachability_global_kernel(const Array<float3> input, const LegDimensions dim, Array<bool> output) {
    lrm_compat::run_reach(input, dim, output);
}
inline void reachability_circles_kernel(const Array<float3> input, const LegDimensions dim, Array<bool> const output) {
    lrm_compat::run_reach(input, dim, output);
}
inline void distance_global_kernel(const Array<float3> input, const LegDimensions dim, Array<float3> output) {
    lrm_compat::run_dist(input, dim, output);
}
inline void distance_circles_kernel(const Array<float3> input, const LegDimensions dim, Array<float3> const output) {
    lrm_compat::run_dist(input, dim, output);
}
inline void forward_kine_kernel(const Array<float3> input, const LegDimensions dim, Array<float3> const output) {
    lrm_compat::run_forward_kine(input, dim, output);
}

// cross_compiled.cuh:4-7 — the reference's template, with the two instantiations it ships
// (cross_compiled.cu:141-151).  Returns the kernel-only milliseconds.
template <typename T_in, typename param, typename T_out>
float apply_kernel(const Array<T_in> input, const param dim,
                   void (*kernel)(const Array<T_in>, const param, Array<T_out> const),
                   Array<T_out> const output);
template <>
inline float apply_kernel<float3, LegDimensions, bool>(const Array<float3> input, const LegDimensions dim,
                                                       void (*kernel)(const Array<float3>, const LegDimensions,
                                                                      Array<bool> const),
                                                       Array<bool> const output) {
    if (kernel == &reachability_global_kernel || kernel == &reachability_circles_kernel)
        return lrm_compat::run_reach(input, dim, output);
    std::fprintf(stderr, "apply_kernel: unknown Array<bool> kernel\n");
    std::exit(EXIT_FAILURE);
}
template <>
inline float apply_kernel<float3, LegDimensions, float3>(const Array<float3> input, const LegDimensions dim,
                                                         void (*kernel)(const Array<float3>, const LegDimensions,
                                                                        Array<float3> const),
                                                         Array<float3> const output) {
    if (kernel == &distance_global_kernel || kernel == &distance_circles_kernel)
        return lrm_compat::run_dist(input, dim, output);
    if (kernel == &forward_kine_kernel) return lrm_compat::run_forward_kine(input, dim, output);
    std::fprintf(stderr, "apply_kernel: unknown Array<float3> kernel\n");
    std::exit(EXIT_FAILURE);
}

// The CPU twins (cross_compiled.cuh:12-15) are DECLARED, so that call sites which pick the compute
// mode at compile time (`if constexpr (ComputeMode == GPUMode) ... else apply_reach_cpu(...)`,
// several_leg.cpp:143-148) keep compiling, and deliberately NOT defined: the product has no CPU
// path, selecting one is a link error.
double apply_reach_cpu(const Array<float3> input, const LegDimensions dim, Array<bool> const output);
double apply_dist_cpu(const Array<float3> input, const LegDimensions dim, Array<float3> const output);

// HeaderCPP.h:54-76 — LegCompact, the reference's unfinished precomputed form of a leg
// (apply_kernel<float3, LegCompact, ...> is instantiated but no kernel takes it, cross_compiled.cu:
// 153-161).  Here it is the public handle of what the library precomputes per (leg, orientation):
// build it once with LegCompacter, pass it to the LegCompact overloads of apply_kernel, and the
// library finds the cached certified tables without rebuilding the plan.
struct LegCompact {
    LegDimensions dim;   // the leg it was built from
    float quat[4];       // body orientation folded into the plan (quatTest = {1,0,0,0} in the reference)
};
inline LegCompact LegCompacter(const LegDimensions dim) {
    LegCompact c;
    c.dim = dim;
    c.quat[0] = 1.f, c.quat[1] = c.quat[2] = c.quat[3] = 0.f;
    return c;
}
inline void reachability_global_kernel(const Array<float3> input, const LegCompact leg, Array<bool> output) {
    float ms = 0.f;
    int rc = lrm_reach(&input.elements->x, input.length, &leg.dim, leg.quat,
                       reinterpret_cast<uint8_t*>(output.elements), 0, nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("reachability kernel (LegCompact)", rc);
}
inline void distance_global_kernel(const Array<float3> input, const LegCompact leg, Array<float3> output) {
    float ms = 0.f;
    int rc = lrm_dist(&input.elements->x, input.length, &leg.dim, leg.quat, &output.elements->x, nullptr, 0,
                      nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("distance kernel (LegCompact)", rc);
}
template <>
inline float apply_kernel<float3, LegCompact, bool>(const Array<float3> input, const LegCompact leg,
                                                    void (*)(const Array<float3>, const LegCompact, Array<bool> const),
                                                    Array<bool> const output) {
    float ms = 0.f;
    int rc = lrm_reach(&input.elements->x, input.length, &leg.dim, leg.quat,
                       reinterpret_cast<uint8_t*>(output.elements), 0, nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("reachability kernel (LegCompact)", rc);
    return ms;
}
template <>
inline float apply_kernel<float3, LegCompact, float3>(const Array<float3> input, const LegCompact leg,
                                                      void (*)(const Array<float3>, const LegCompact,
                                                               Array<float3> const),
                                                      Array<float3> const output) {
    float ms = 0.f;
    int rc = lrm_dist(&input.elements->x, input.length, &leg.dim, leg.quat, &output.elements->x, nullptr, 0,
                      nullptr, &ms);
    if (rc != LRM_OK) lrm_compat::die("distance kernel (LegCompact)", rc);
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
