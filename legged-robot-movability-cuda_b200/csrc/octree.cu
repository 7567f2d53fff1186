// octree.cu — body-space octree positionability, replaces apply_oct / branchKernel /
// validity_child (several_leg_octree.cu:19-151,241-377,391-488) and the pointer tree helpers of
// octree_util.cu (copyTreeOnCpu, countLeaf, extractValidAsArray).
//
// The reference grows a pointer-linked tree on the device heap with dynamic parallelism (one
// kernel per node, device cudaMalloc, unsynchronised shared flags) and evaluates every
// (child, foothold, orientation sample) triple of a node in one flat launch.  Here the tree is a
// flat host-side array grown level by level; per refinement pass ONE kernel evaluates all children
// of all nodes expanded in that pass: a CTA slice per child walks only the foothold cells inside
// the child's elongated box (cell grid of cell_grid.cuh), every lane takes a foothold and runs the
// (orientation sample x 4 legs) distance evaluations from per-(sample, leg) plans built on the
// host, and the three per-child flags are OR-reduced with warp votes and one atomicOr.  Results
// follow the sequential semantics (flags OR-ed over all work items of a pass, applied after the
// pass); the reference's own GPU outcome depends on a shared-memory race.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "cell_grid.cuh"
#include "kernels.h"
#include "leg_math.cuh"

namespace lrm {

namespace {

constexpr float kMinBox = 100.f;       // MINBOXSIZE, settings.h:17
constexpr int kDeadQuadrant = 128;     // DEADQUADRAN, settings.h:30
constexpr float kRotBelow = 50.f;      // EnableRotBelow, settings.h:32
constexpr float kConvexRadius = 100.f; // convexRadius, settings.h:33
constexpr float kRootHalf = 5000.f;    // BoxSize, settings.h:26
constexpr int kLegs = 4;               // LegCount, settings.h:41
constexpr int kSamples = 27;           // AngleSample 3x3x3, settings.h:34

struct Box {
    float c[3], h[3];
};

struct HostNode {
    Box box;
    bool validity = false, leaf = false, raw = false, on_edge = false;
    int children = -1;  // index of the first of 8 consecutive children, -1 when none
};

bool null_box(const Box& b) {
    for (int i = 0; i < 3; i++)
        if (b.c[i] != 0.f || b.h[i] != 0.f) return false;
    return true;
}

// CreateChildBox (octree_util.cu.h:105-151) for SUB_QUAD = 1, quadCount = 3, no "small" axes
// (branchKernel always passes {0,0,0}, several_leg_octree.cu:258,324).  The child index is
// bit-reversed; an axis whose half extent is already below MINBOXSIZE is not split, and the
// children that would duplicate along it are dead.
int child_box(const Box& parent, unsigned index, Box* child) {
    *child = parent;
    unsigned sub = ((index & 1u) << 2) | (index & 2u) | ((index & 4u) >> 2);  // reverse 3 bits
    float div[3] = {2.f, 2.f, 2.f};
    int missing = 0;
    for (unsigned q = 0; q < 3; q++) {
        if (child->h[q] < kMinBox) {
            missing++;
            if ((sub >> 2) & 1u) return kDeadQuadrant;
            const unsigned mask = 0xfu ^ ((1u << q) - 1u);  // bitShiftBetween(sub, q, 3, 1)
            sub = (sub & ~mask) | (((sub & mask) << 1) & mask);
            div[q] = 1.f;
        }
    }
    for (int q = 0; q < 3; q++) {
        const float old = child->h[q];
        child->h[q] = child->h[q] / div[q];
        float mv = old - child->h[q];
        if ((sub >> q) & 1u) mv *= -1.f;  // flipVectorOnQuad
        child->c[q] += mv;
    }
    return missing;
}

// ---- device side -------------------------------------------------------------------------------
struct ChildTask {
    Box box;            // the child
    float elong[3];     // parent half extents + leg reach (the pre-filter box, :76-82)
    float margin;       // 0 when the orientation samples are active, else EnableRotBelow / 3
    int n_samples;      // 27 when parent half extent < EnableRotBelow, else 1
    int parent_valid;
    int skip;           // child already valid before the pass (dead quadrant or processed)
    int wedge;          // the coxa yaw limits span less than pi (cell / foothold pruning)
};

__device__ __forceinline__ bool in_box(float x, float y, float z, const float* h) {
    const float ex = fabsf(h[0]), ey = fabsf(h[1]), ez = fabsf(h[2]);  // isInBox, octree_util.cu.h:153-159
    return ex >= x && ey >= y && ez >= z && -ex < x && -ey < y && -ez < z;
}

// flags per child: bit 0 = some work item was on the edge, bit 1 = some was reachable,
// bit 2 = some was reachable and not on the edge (valid leaf)
__global__ void __launch_bounds__(256) oct_validity_kernel(const ChildTask* __restrict__ tasks, CellGrid g,
                                                           const LegPlan* __restrict__ plans,
                                                           unsigned* __restrict__ flags, int slices) {
    const ChildTask T = tasks[blockIdx.y];
    if (T.skip) return;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float edge_raw = T.box.h[0] * T.box.h[0] + T.box.h[1] * T.box.h[1] + T.box.h[2] * T.box.h[2];
    const bool by_box = edge_raw > kConvexRadius * kConvexRadius;
    // A leg matters for a foothold only if it reaches it or its distance vector lands inside the
    // child (|d| at most the child's half diagonal, or the convex radius rule): a foothold farther
    // than that from the leg's workspace skips the leg's distance evaluation — exactly, the
    // returned vector always ends on the workspace boundary, so it is at least that long.
    const float far = sqrtf(edge_raw + T.margin) * 1.001f + 0.05f;

    // cells overlapping the elongated box around the child centre
    const int cx0 = max((int)floorf((T.box.c[0] - T.elong[0] - g.x0) * g.inv_cell), 0);
    const int cx1 = min((int)floorf((T.box.c[0] + T.elong[0] - g.x0) * g.inv_cell), g.nx - 1);
    const int cy0 = max((int)floorf((T.box.c[1] - T.elong[1] - g.y0) * g.inv_cell), 0);
    const int cy1 = min((int)floorf((T.box.c[1] + T.elong[1] - g.y0) * g.inv_cell), g.ny - 1);
    const int w = cx1 - cx0 + 1, h = cy1 - cy0 + 1;
    unsigned mine = 0;
    if (w > 0 && h > 0) {
        const int ncell = w * h;
        // (slice, warp) pairs stride over the cells; lanes over the footholds of a cell
        for (int k = blockIdx.x * nwarp + warp; k < ncell; k += slices * nwarp) {
            const int ccx = cx0 + k % w, ccy = cy0 + k / w;
            const int c = ccy * g.nx + ccx;
            const float2 zr = g.cell_z[c];
            if (zr.x > T.box.c[2] + T.elong[2] || zr.y < T.box.c[2] - T.elong[2]) continue;
            // the same question for the whole cell (its bounding ball): lanes 0..n_samples*4-1 take
            // one (sample, leg) plan each; a cell no plan can care about is not even read
            {
                const float2 ball = g.cell_ball[c];
                if (ball.y < 0.f) continue;
                const float cell = 1.0f / g.inv_cell;
                const float bx = g.x0 + ((float)ccx + 0.5f) * cell - T.box.c[0];
                const float by = g.y0 + ((float)ccy + 0.5f) * cell - T.box.c[1];
                const float bz = ball.x - T.box.c[2];
                bool any = false;
                for (int q = lane; q < T.n_samples * kLegs; q += 32) {
                    const LegPlan& L = plans[q];
                    any = any || leg_ball_possible(L, to_coxa_frame(L, bx, by, bz), far + ball.y, T.wedge != 0);
                }
                // a valid parent marks the child for ANY foothold inside the elongated box (bit 1/2
                // of the flags do not depend on the legs then): such cells must be read
                if (!__any_sync(0xffffffffu, any) && !T.parent_valid) continue;
            }
            const int beg = g.cell_start[c], end = g.cell_start[c + 1];
            for (int i = beg + lane; i < end; i += 32) {
                const float4 f = g.pts[i];
                const float vx = f.x - T.box.c[0], vy = f.y - T.box.c[1], vz = f.z - T.box.c[2];
                if (!in_box(vx, vy, vz, T.elong)) continue;
                for (int a = 0; a < T.n_samples; a++) {
                    int reach = 0, cross = 0;
#pragma unroll 1
                    for (int leg = 0; leg < kLegs; leg++) {
                        const LegPlan& L = plans[a * kLegs + leg];
                        const SectorTable& tab = *reinterpret_cast<const SectorTable*>(&L.sector[0]);
                        const CoxaPoint pc = to_coxa_frame(L, vx, vy, vz);
                        if (!leg_ball_possible(L, pc, far, T.wedge != 0)) continue;
                        // distance() = distance_global (one_leg_global.cu:253-264)
                        const DistResult d = L.generic ? dist_coxa_frame<true>(L, tab, pc)
                                                       : dist_coxa_frame<false>(L, tab, pc);
                        reach += d.flag ? 1 : 0;
                        const bool in = by_box ? in_box(d.dx, d.dy, d.dz, T.box.h)  // :99-103
                                               : (d.dx * d.dx + d.dy * d.dy + d.dz * d.dz) < edge_raw + T.margin;
                        cross += in ? 1 : 0;
                    }
                    const bool edge = cross > 0;  // LegCount - LegNumberForStab = 0
                    const bool ok = T.parent_valid || reach >= kLegs;
                    mine |= (edge ? 1u : 0u) | (ok ? 2u : 0u) | ((ok && !edge) ? 4u : 0u);
                }
            }
        }
    }
    // OR over the warp, then one atomic per warp
    for (int o = 16; o; o >>= 1) mine |= __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0 && mine) atomicOr(&flags[blockIdx.y], mine);
}

// plans for the 27 orientation samples x 4 leg mounts (octree_util.cu.h:184-198, settings.h:41)
void build_oct_plans(const lrm_leg_t& leg, std::vector<LegPlan>* out) {
    const float pi = 3.14159265358979323846264338327950288419716939937510582097f;
    std::vector<LegPlan>& plans = *out;
    plans.assign((size_t)kSamples * kLegs, LegPlan{});
    for (int a = 0; a < kSamples; a++) {
        // QuaternionFromAngleIndex: (ind + ind/2) % 3 maps 2 -> 0, so only {min, mid} are sampled
        const float lim[6] = {-pi / 4, pi / 4, -pi / 8, pi / 8, -pi / 8, pi / 8};
        float rpy[3];
        unsigned rest = (unsigned)a;
        for (int i = 0; i < 3; i++) {
            unsigned char ind = (unsigned char)(rest % 3);
            ind = (unsigned char)((ind + (ind / 2)) % 3);
            rest /= 3;
            const float x = (float)ind / (unsigned char)2;
            rpy[i] = (1 - x) * lim[i * 2] + x * lim[i * 2 + 1];
        }
        const float ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
        float qr[4], qp[4], qy[4], tmp[4];
        quat_from_vect_angle(ax, rpy[0], qr);
        quat_from_vect_angle(ay, rpy[1], tmp);
        quat_multiply(tmp, qr, qp);
        quat_from_vect_angle(az, rpy[2], tmp);
        quat_multiply(tmp, qp, qy);  // RPYtoQuat, octree_util.cu.h:164-172
        for (int k = 0; k < kLegs; k++) {
            lrm_leg_t l = leg;
            l.body_angle = pi / 4 * k;
            build_leg_plan(l, qy, &plans[(size_t)a * kLegs + k]);
        }
    }
}

// The 8 children of `parent` as branchKernel initialises them (several_leg_octree.cu:315-352),
// with the work descriptor validity_child needs for each.
void init_children(const HostNode& parent, const lrm_leg_t& leg, HostNode* children, std::vector<ChildTask>* tasks) {
    const float reach = leg.body + leg.coxa_length + leg.femur_length + leg.tibia_length;
    const bool rot = parent.box.h[0] < kRotBelow;
    for (unsigned i = 0; i < 8; i++) {
        HostNode& ch = children[i];
        Box nb;
        const int missing = child_box(parent.box, i, &nb);
        ChildTask t{};
        if (missing == kDeadQuadrant) {  // :331-339
            ch.leaf = true, ch.raw = false, ch.validity = true, ch.on_edge = true;
            ch.box = Box{{0, 0, 0}, {0, 0, 0}};
            t.skip = 1;
        } else {
            ch.on_edge = false, ch.validity = false, ch.box = nb;
            ch.leaf = (3 - missing) <= 0;
            ch.raw = !ch.leaf;
        }
        t.box = ch.box;
        for (int q = 0; q < 3; q++) t.elong[q] = parent.box.h[q] + reach;
        t.margin = rot ? 0.f : kRotBelow / 3;
        t.n_samples = rot ? kSamples : 1;
        t.parent_valid = parent.validity ? 1 : 0;
        t.wedge = (leg.max_angle_coxa >= leg.min_angle_coxa && leg.max_angle_coxa - leg.min_angle_coxa < 3.0f) ? 1 : 0;
        tasks->push_back(t);
    }
}

// validity_child's write-back (:134-150) from the OR-ed flags of a pass
void apply_flags(unsigned f, HostNode* ch) {
    if (f & 2u) ch->validity = true;
    if (f & 4u) ch->leaf = true;
    if ((f & 1u) && !(f & 4u)) ch->on_edge = true;
}

// One kernel pass over `tasks`; flags come back on the host.
cudaError_t evaluate_tasks(const std::vector<ChildTask>& tasks, const CellGrid& grid, const LegPlan* d_plans,
                           cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1, std::vector<unsigned>* flags) {
    ChildTask* d_tasks = nullptr;
    unsigned* d_flags = nullptr;
    flags->assign(tasks.size(), 0u);
    cudaError_t status = cudaMalloc((void**)&d_tasks, tasks.size() * sizeof(ChildTask));
    if (status != cudaSuccess) return status;
    status = cudaMalloc((void**)&d_flags, tasks.size() * sizeof(unsigned));
    if (status == cudaSuccess) {
        cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * sizeof(ChildTask), cudaMemcpyHostToDevice, stream);
        cudaMemsetAsync(d_flags, 0, tasks.size() * sizeof(unsigned), stream);
        // enough slices that a pass with few children (the first ones) still fills the GPU
        int slices = (int)(148 * 4 / tasks.size()) + 1;
        if (slices > 64) slices = 64;
        if (ev0) cudaEventRecord(ev0, stream);
        for (size_t off = 0; off < tasks.size(); off += 32768) {
            const unsigned cnt = (unsigned)std::min<size_t>(32768, tasks.size() - off);
            oct_validity_kernel<<<dim3((unsigned)slices, cnt), 256, 0, stream>>>(d_tasks + off, grid, d_plans,
                                                                                d_flags + off, slices);
        }
        status = cudaGetLastError();
        if (ev1) cudaEventRecord(ev1, stream);
        if (status == cudaSuccess)
            status = cudaMemcpyAsync(flags->data(), d_flags, flags->size() * sizeof(unsigned), cudaMemcpyDeviceToHost,
                                     stream);
        if (status == cudaSuccess) status = cudaStreamSynchronize(stream);
    }
    cudaFree(d_tasks);
    cudaFree(d_flags);
    return status;
}

#define OCT_CHECK(call)                    \
    do {                                   \
        cudaError_t e_ = (call);           \
        if (e_ != cudaSuccess) return e_;  \
    } while (0)

}  // namespace

cudaError_t run_octree(const float* d_footholds, size_t nt, const lrm_leg_t& leg, int max_depth,
                       std::vector<float>* centres, cudaStream_t stream, float* kernel_ms, int shard, int nshards,
                       size_t* child_counts) {
    centres->clear();
    if (child_counts)
        for (int i = 0; i < 8; i++) child_counts[i] = 0;
    if (nshards < 1 || shard < 0 || shard >= nshards) return cudaErrorInvalidValue;
    if (kernel_ms) *kernel_ms = 0.f;
    if (nt > 0x7fffffffull) return cudaErrorInvalidValue;
    DevBuf mem;
    CellGrid grid;
    OCT_CHECK(build_grid(mem, d_footholds, nt, nullptr, 0.f, stream, &grid));

    std::vector<LegPlan> plans;
    build_oct_plans(leg, &plans);
    LegPlan* d_plans;
    OCT_CHECK(mem.alloc(&d_plans, plans.size()));
    OCT_CHECK(cudaMemcpyAsync(d_plans, plans.data(), plans.size() * sizeof(LegPlan), cudaMemcpyHostToDevice, stream));

    std::vector<HostNode> nodes(1);
    nodes[0].box = Box{{0.f, 0.f, 0.f}, {kRootHalf, kRootHalf, kRootHalf}};
    nodes[0].raw = true;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (kernel_ms) {
        OCT_CHECK(cudaEventCreate(&ev0));
        OCT_CHECK(cudaEventCreate(&ev1));
    }
    cudaError_t status = cudaSuccess;
    for (int pass = 0; pass < max_depth && status == cudaSuccess; pass++) {
        // branchKernel's descent (:241-313): find the raw nodes this pass expands
        std::vector<int> expand, stack{0};
        while (!stack.empty()) {
            const int n = stack.back();
            stack.pop_back();
            if (nodes[n].raw) {
                expand.push_back(n);
                continue;
            }
            if (nodes[n].children < 0) continue;
            for (int i = 7; i >= 0; i--) {
                HostNode& ch = nodes[nodes[n].children + i];
                if (!ch.on_edge) ch.leaf = true;
                if (!ch.leaf) stack.push_back(nodes[n].children + i);
            }
        }
        if (expand.empty()) break;
        std::vector<ChildTask> tasks;
        tasks.reserve(expand.size() * 8);
        for (int n : expand) {
            const int first = (int)nodes.size();
            nodes.resize(nodes.size() + 8);
            nodes[n].children = first;
            init_children(nodes[n], leg, &nodes[first], &tasks);
            nodes[n].raw = false;
            if (n == 0 && nshards > 1) {
                // sharding by top-level children (dealt round-robin): the subtrees under the root's
                // children never interact, so a shard simply never looks at the other shards' children
                for (int i = 0; i < 8; i++)
                    if (i % nshards != shard) {
                        HostNode& ch = nodes[first + i];
                        tasks[tasks.size() - 8 + i].skip = 1;
                        ch.leaf = true, ch.raw = false, ch.validity = false, ch.on_edge = false;
                    }
            }
        }
        std::vector<unsigned> flags;
        status = evaluate_tasks(tasks, grid, d_plans, stream, ev0, ev1, &flags);
        if (status == cudaSuccess && kernel_ms) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) *kernel_ms += ms;
        }
        if (status == cudaSuccess) {
            size_t t = 0;
            for (int n : expand)
                for (int i = 0; i < 8; i++, t++)
                    if (!tasks[t].skip) apply_flags(flags[t], &nodes[nodes[n].children + i]);
        }
    }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (status != cudaSuccess) return status;

    // fill_recus (octree_util.cu:123-147): depth-first, children in index order
    if (nodes[0].children >= 0) {
        std::vector<std::pair<int, int>> stack{{0, 0}};
        while (!stack.empty()) {
            auto& top = stack.back();
            if (top.second == 8) {
                stack.pop_back();
                continue;
            }
            const HostNode& ch = nodes[nodes[top.first].children + top.second];
            top.second++;
            const bool endpoint = !(ch.leaf || ch.raw || null_box(ch.box));
            const bool valid = !null_box(ch.box) && (ch.leaf || ch.raw) && ch.validity;
            if (endpoint && ch.children >= 0) {
                stack.push_back({(int)(&ch - nodes.data()), 0});
            } else if (valid) {
                if (child_counts) child_counts[stack.size() == 1 ? top.second - 1 : stack[1].first - nodes[0].children]++;
                centres->push_back(ch.box.c[0]);
                centres->push_back(ch.box.c[1]);
                centres->push_back(ch.box.c[2]);
            }
        }
    }
    return cudaSuccess;
}

// validity_child on the 8 children of one parent box (lrm_oct_children).
cudaError_t run_octree_children(const float* d_footholds, size_t nt, const lrm_leg_t& leg, const float* parent_box6,
                                int parent_validity, uint8_t* flags32, float* boxes48, cudaStream_t stream) {
    if (nt > 0x7fffffffull) return cudaErrorInvalidValue;
    DevBuf mem;
    CellGrid grid;
    OCT_CHECK(build_grid(mem, d_footholds, nt, nullptr, 0.f, stream, &grid));
    std::vector<LegPlan> plans;
    build_oct_plans(leg, &plans);
    LegPlan* d_plans;
    OCT_CHECK(mem.alloc(&d_plans, plans.size()));
    OCT_CHECK(cudaMemcpyAsync(d_plans, plans.data(), plans.size() * sizeof(LegPlan), cudaMemcpyHostToDevice, stream));
    HostNode parent;
    std::memcpy(&parent.box, parent_box6, sizeof(Box));
    parent.validity = parent_validity != 0;
    parent.raw = true;
    HostNode children[8];
    std::vector<ChildTask> tasks;
    init_children(parent, leg, children, &tasks);
    std::vector<unsigned> flags;
    OCT_CHECK(evaluate_tasks(tasks, grid, d_plans, stream, nullptr, nullptr, &flags));
    for (int i = 0; i < 8; i++) {
        if (!tasks[i].skip) apply_flags(flags[i], &children[i]);
        flags32[4 * i + 0] = children[i].validity, flags32[4 * i + 1] = children[i].leaf;
        flags32[4 * i + 2] = children[i].raw, flags32[4 * i + 3] = children[i].on_edge;
        std::memcpy(boxes48 + 6 * i, &children[i].box, sizeof(Box));
    }
    return cudaSuccess;
}

}  // namespace lrm
