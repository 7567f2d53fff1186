// Call-site shapes of the reference, against include/lrm_compat.hpp (compile-only on the CPU suite,
// run on the GPU suite).  Each block mirrors how a reference caller spells the call:
//   bench.cpp:120-152          apply_kernel(target_map, dim, reachability_global_kernel, out)
//   several_leg.cpp:143-148    if constexpr (ComputeMode == GPUMode) ... else apply_reach_cpu(...)
//   cross_compiled.cu:141-151  the explicit instantiations apply_kernel<float3, LegDimensions, bool>
//   (generic)                  the kernel kept in a function-pointer variable
//   HeaderCPP.h:54-76          LegCompact through the same template
#include <cstdio>
#include <cstring>
#include <vector>

#include "lrm_compat.hpp"

enum { GPUMode, CPUMode };
constexpr int ComputeMode = GPUMode;

int main() {
    LegDimensions dim = get_M2_leg(0);
    std::vector<float3> pts;
    for (float x = -100; x <= 601; x += 7.f)
        for (float y = -40; y <= 40; y += 40.f)
            for (float z = -300; z <= 51; z += 7.f) pts.push_back({x, y, z});
    Array<float3> target_map{pts.size(), pts.data()};

    Array<bool> out2;
    out2.length = target_map.length;
    out2.elements = new bool[out2.length];
    float duration;
    if constexpr (ComputeMode == GPUMode)
        duration = apply_kernel(target_map, dim, reachability_global_kernel, out2);
    else
        duration = apply_reach_cpu(target_map, dim, out2);   // declared, never defined: no CPU path

    // explicit template arguments, as the reference instantiates them
    Array<bool> out3{target_map.length, new bool[target_map.length]};
    duration += apply_kernel<float3, LegDimensions, bool>(target_map, dim, reachability_circles_kernel, out3);
    if (std::memcmp(out2.elements, out3.elements, out2.length) != 0) return 2;

    // the kernel in a function-pointer variable
    void (*dist_kernel)(const Array<float3>, const LegDimensions, Array<float3>) = distance_global_kernel;
    Array<float3> d1{target_map.length, new float3[target_map.length]};
    Array<float3> d2{target_map.length, new float3[target_map.length]};
    duration += apply_kernel(target_map, dim, dist_kernel, d1);
    duration += apply_kernel<float3, LegDimensions, float3>(target_map, dim, distance_circles_kernel, d2);
    if (std::memcmp(d1.elements, d2.elements, d1.length * sizeof(float3)) != 0) return 3;

    // a handle called directly runs the same sweep
    Array<float3> d3{target_map.length, new float3[target_map.length]};
    distance_global_kernel(target_map, dim, d3);
    if (std::memcmp(d1.elements, d3.elements, d1.length * sizeof(float3)) != 0) return 4;

    // LegCompact (precomputed leg): same results through the same template
    LegCompact compact = LegCompacter(dim);
    Array<float3> d4{target_map.length, new float3[target_map.length]};
    void (*compact_kernel)(const Array<float3>, const LegCompact, Array<float3>) = distance_global_kernel;
    duration += apply_kernel(target_map, compact, compact_kernel, d4);
    if (std::memcmp(d1.elements, d4.elements, d1.length * sizeof(float3)) != 0) return 5;

    size_t reachable = 0;
    for (size_t i = 0; i < out2.length; i++) reachable += out2.elements[i];
    std::printf("call sites: %zu of %zu reachable, %.4f ms\n", reachable, out2.length, duration);
    delete[] out2.elements, delete[] out3.elements;
    delete[] d1.elements, delete[] d2.elements, delete[] d3.elements, delete[] d4.elements;
    return reachable > 0 ? 0 : 1;
}
