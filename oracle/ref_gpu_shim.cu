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
#include "octree_util.cu.h"
#include "settings.h"

#include <cstdint>
#include <cstring>

// defined in several_leg_octree.cu:19 (no header declares it)
__global__ void validity_child(Node parent, const Array<float3> input, const LegDimensions leg);

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

// validity_child (several_leg_octree.cu:19-151) on a hand-built one-level tree: the per-child
// predicate of apply_oct WITHOUT the recursion, the device heap and the dynamic parallelism that
// keep apply_oct itself from terminating on sm_100.  The eight children are initialised the way
// branchKernel / branchCpu do it (several_leg_octree.cu:168-199,315-352) with the reference's own
// CreateChildBox; then the reference's kernel runs on them, unmodified.
// The kernel runs as ONE thread (<<<1, 1>>>).  That is the only launch shape in which it is
// well defined: distance() -> distance_global (one_leg_global.cu:74-101) stages the oriented leg in
// a __shared__ variable written by threadIdx.x == 0 and then calls __syncthreads() — inside
// validity_child's divergent loop (`continue`s per work item, :52-60, :82), so with more than one
// thread per block the barrier sits in divergent code (it deadlocked the B200 for 25 minutes;
// this is also why apply_oct never terminates on sm_100) and every thread would use thread 0's
// orientation sample.  One thread = the sequential semantics, deterministic.
// out_flags: 8 x {validity, leaf, raw, onEdge}; out_boxes: 8 x {center xyz, topOffset xyz}.
int refgpu_validity_child(const float* parent_box6, int parent_validity, const float* footholds,
                          size_t nt, const float* leg14, uint8_t* out_flags, float* out_boxes) {
    Node parent;
    std::memcpy(&parent.box, parent_box6, sizeof(Box));
    parent.validity = parent_validity != 0;
    parent.leaf = false, parent.raw = true, parent.onEdge = false;
    parent.childrenCount = MaxChildQuad;
    if (cudaMallocManaged(&parent.childrenArr, MaxChildQuad * sizeof(Node)) != cudaSuccess) return -1;
    const bool small[3] = {0, 0, 0};
    for (uint c = 0; c < MaxChildQuad; c++) {
        Node& node = parent.childrenArr[c];
        node = Node();
        Box nb;
        uchar missing;
        CreateChildBox(parent.box, nb, 3, c, small, missing);
        node.childrenCount = MaxChildQuad;
        node.childrenArr = nullptr;
        if (missing == DEADQUADRAN) {
            node.leaf = true, node.raw = false, node.validity = true, node.onEdge = true;
            node.box = NullBox;
            continue;
        }
        node.onEdge = false, node.validity = false, node.box = nb;
        node.leaf = (3 - missing) <= 0;
        node.raw = !node.leaf;
    }
    Array<float3> in{nt, nullptr};
    if (cudaMalloc(&in.elements, (nt ? nt : 1) * sizeof(float3)) != cudaSuccess) return -2;
    cudaMemcpy(in.elements, footholds, nt * sizeof(float3), cudaMemcpyHostToDevice);
    validity_child<<<1, 1>>>(parent, in, leg_from(leg14));
    const cudaError_t e = cudaDeviceSynchronize();
    for (uint c = 0; c < MaxChildQuad; c++) {
        const Node& node = parent.childrenArr[c];
        out_flags[4 * c + 0] = node.validity, out_flags[4 * c + 1] = node.leaf;
        out_flags[4 * c + 2] = node.raw, out_flags[4 * c + 3] = node.onEdge;
        std::memcpy(out_boxes + 6 * c, &node.box, sizeof(Box));
    }
    cudaFree(in.elements);
    cudaFree(parent.childrenArr);
    return e == cudaSuccess ? 0 : -3;
}

}  // extern "C"
