// Compile-only check (CPU suite) + GPU run (gpu suite) that a reference-style call site builds
// against include/lrm_compat.hpp: the body below is the shape of bench.cpp:120-152.
#include <cstdio>
#include <vector>

#include "lrm_compat.hpp"

int main() {
    LegDimensions dim = get_M2_leg(0);
    std::vector<float3> pts;
    for (float x = -100; x <= 601; x += 5.f)
        for (float z = -100; z <= 51; z += 5.f) pts.push_back({x, 0.f, z});
    Array<float3> target_map{pts.size(), pts.data()};

    Array<bool> out;
    out.length = target_map.length;
    out.elements = new bool[out.length];
    float duration = apply_kernel(target_map, dim, reachability_global_kernel, out);
    size_t reachable = 0;
    for (size_t i = 0; i < out.length; i++) reachable += out.elements[i];
    std::printf("reach: %zu of %zu reachable, %.4f ms\n", reachable, out.length, duration);
    delete[] out.elements;

    Array<float3> dist;
    dist.length = target_map.length;
    dist.elements = new float3[dist.length];
    duration = apply_kernel(target_map, dim, distance_global_kernel, dist);
    std::printf("dist: d[0] = (%.4f, %.4f, %.4f), %.4f ms\n", dist.elements[0].x, dist.elements[0].y,
                dist.elements[0].z, duration);
    delete[] dist.elements;
    return reachable > 0 ? 0 : 1;
}
