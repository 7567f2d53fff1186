// TEST / BENCH INFRASTRUCTURE ONLY — never linked into the product library.
//
// extern "C" window onto robot_full_struct (several_leg.cu:796-877), the reference's brute-force
// 4-leg positionability pipeline.  several_leg.cu is commented out of the reference's CMake at
// HEAD; it compiles unmodified with `-include settings.h -include one_leg.cu`, which
// is how oracle/Makefile builds it into oracle/_ref/libref_gpu_several.so (a separate library,
// because that forced include defines the one_leg.cu kernels a second time).
#include "HeaderCPP.h"
#include "HeaderCUDA.h"
#include "several_leg.cu.h"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <tuple>

extern "C" {

// bodies (nb x 3), map (nt x 3), legs (nlegs x 14 floats; the reference reads the first 4).
// out_xyz receives the compacted list of standable body positions (up to cap); returns wall ms.
double refgpu_full_struct(const float* bodies, size_t nb, const float* map, size_t nt,
                          const float* legs14, size_t nlegs, float* out_xyz, size_t cap,
                          size_t* count) {
    static_assert(sizeof(LegDimensions) == 14 * sizeof(float), "LegDimensions layout");
    Array<float3> b{nb, (float3*)bodies};
    Array<float3> m{nt, (float3*)map};
    Array<LegDimensions> l{nlegs, (LegDimensions*)legs14};
    const auto t0 = std::chrono::steady_clock::now();
    auto out = robot_full_struct(b, m, l);
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    Array<float3> ob = std::get<0>(out);
    Array<int> oc = std::get<1>(out);
    *count = ob.length;
    const size_t k = ob.length < cap ? ob.length : cap;
    if (k) std::memcpy(out_xyz, ob.elements, k * sizeof(float3));
    delete[] ob.elements;
    delete[] oc.elements;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

}  // extern "C"
