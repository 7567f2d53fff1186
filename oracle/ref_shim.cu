// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// extern "C" window onto the UNMODIFIED reference sources under /root/reference.
// This translation unit #includes the reference's own one_leg_global.cu (which
// pulls one_leg.cu, circles.cu.h, leg_geometry.cu.h, unified_math_cuda.cu.h,
// octree_util.cu.h, settings.h) where it lies; nothing is copied into this repo.
// Every function below only forwards to a reference host function, so the
// resulting library (oracle/_ref/libref_oracle.so) *is* the reference CPU path
// (one_leg_global.cu:132-147) plus the host-callable primitives the CPU
// restatement (oracle/oracle_port.c) is pinned against.
//
// Build: see oracle/Makefile (nvcc host pass, -O3, no fast-math on host — same as
// the reference's CMakeLists.txt:145-148).  Only usable where /root/reference
// exists (this container); the GPU box receives the prebuilt .so.
#include "settings.h"          // must precede circles.cu.h (MegaClamp)
#include "one_leg_global.cu"   // reference TU: reachability_global, distance_global, *_kernel_cpu
#include "static_variables.h"  // leg_factory / get_M2_leg / get_moonbot_leg

#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {
inline LegDimensions leg_from(const float* l) {
    LegDimensions d;
    static_assert(sizeof(LegDimensions) == 14 * sizeof(float), "LegDimensions layout");
    std::memcpy(&d, l, sizeof(d));
    return d;
}
inline Quaternion quat_from(const float* q) { return make_float4(q[0], q[1], q[2], q[3]); }

template <class F> void parallel_slabs(size_t n, int threads, F f) {
    if (threads <= 1 || n < 4096) {
        f(0, n);
        return;
    }
    std::vector<std::thread> pool;
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        pool.emplace_back([=] { f(lo, hi); });
    }
    for (auto& th : pool) th.join();
}
} // namespace

extern "C" {

// static_variables.cpp:44-93
void ref_get_leg(int robot, float azimuth, float* out14) {
    LegDimensions d = (robot == 0) ? get_moonbot_leg(azimuth) : get_M2_leg(azimuth);
    std::memcpy(out14, &d, sizeof(d));
}

// one_leg_global.cu:106-130 (host branch) over a slab, arbitrary body quaternion.
void ref_reach(const float* xyz, size_t n, const float* leg14, const float* quat4, uint8_t* out,
               int threads) {
    const LegDimensions dim = leg_from(leg14);
    const Quaternion q = quat_from(quat4);
    const float3* p = reinterpret_cast<const float3*>(xyz);
    parallel_slabs(n, threads, [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) out[i] = reachability_global(p[i], dim, q) ? 1 : 0;
    });
}

// one_leg_global.cu:74-101 (host branch): vector to the reachability edge + flag.
void ref_dist(const float* xyz, size_t n, const float* leg14, const float* quat4, float* out_xyz,
              uint8_t* out_flag, int threads) {
    const LegDimensions dim = leg_from(leg14);
    const Quaternion q = quat_from(quat4);
    const float3* p = reinterpret_cast<const float3*>(xyz);
    float3* o = reinterpret_cast<float3*>(out_xyz);
    parallel_slabs(n, threads, [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            float3 r = p[i];
            bool f = distance_global(r, dim, q);
            o[i] = r;
            if (out_flag) out_flag[i] = f ? 1 : 0;
        }
    });
}

// one_leg.cu:280-341 without the body-orientation wrapper.
void ref_reach_circles(const float* xyz, size_t n, const float* leg14, uint8_t* out) {
    const LegDimensions dim = leg_from(leg14);
    const float3* p = reinterpret_cast<const float3*>(xyz);
    for (size_t i = 0; i < n; i++) out[i] = reachability_circles(p[i], dim) ? 1 : 0;
}
void ref_dist_circles(const float* xyz, size_t n, const float* leg14, float* out_xyz,
                      uint8_t* out_flag) {
    const LegDimensions dim = leg_from(leg14);
    const float3* p = reinterpret_cast<const float3*>(xyz);
    float3* o = reinterpret_cast<float3*>(out_xyz);
    for (size_t i = 0; i < n; i++) {
        float3 r = p[i];
        bool f = distance_circles(r, dim);
        o[i] = r;
        if (out_flag) out_flag[i] = f ? 1 : 0;
    }
}

// The CPU path exactly as shipped (single thread, quatTest): one_leg_global.cu:132-147,
// timed like cross_compiled.cu:163-181.  Returns milliseconds.
double ref_apply_reach_cpu(const float* xyz, size_t n, const float* leg14, uint8_t* out) {
    Array<float3> in{n, const_cast<float3*>(reinterpret_cast<const float3*>(xyz))};
    Array<bool> o{n, reinterpret_cast<bool*>(out)};
    auto t0 = std::chrono::high_resolution_clock::now();
    reachability_kernel_cpu(in, leg_from(leg14), o);
    auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}
double ref_apply_dist_cpu(const float* xyz, size_t n, const float* leg14, float* out_xyz) {
    Array<float3> in{n, const_cast<float3*>(reinterpret_cast<const float3*>(xyz))};
    Array<float3> o{n, reinterpret_cast<float3*>(out_xyz)};
    auto t0 = std::chrono::high_resolution_clock::now();
    distance_kernel_cpu(in, leg_from(leg14), o);
    auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// circles.cu.h:48-78 -> bit0 UpperRegion, bit1 FullyExtended, bit2 FemurAngleLimitation,
// bit3 FemurAngleLimitation_other
int ref_find_region(float x, float y, const float* leg14) {
    AreaBitField r = find_region(x, y, leg_from(leg14));
    return (int)r.UpperRegion | ((int)r.FullyExtended << 1) | ((int)r.FemurAngleLimitation << 2) |
           ((int)r.FemurAngleLimitation_other << 3);
}
// circles.cu.h:337-383 at the region of (x, y): out = 4 x {cx, cy, r, attractive}
int ref_insert_circles(float x, float y, const float* leg14, float* out16) {
    const LegDimensions d = leg_from(leg14);
    AreaBitField r = find_region(x, y, d);
    Circle c[MAX_CIRCLES];
    Circle* tail = insert_circles(d, r, c);
    int n = (int)(tail - c);
    for (int i = 0; i < n; i++) {
        out16[4 * i + 0] = c[i].x;
        out16[4 * i + 1] = c[i].y;
        out16[4 * i + 2] = c[i].radius;
        out16[4 * i + 3] = c[i].attractivity ? 1.f : 0.f;
    }
    return n;
}
// circles.cu.h:417-476: out = up to 10 x {x, y}
int ref_insert_intersec(const float* leg14, float* out20) {
    Circle c[MAX_INTERSECT];
    Circle* tail = insert_intersecv2(leg_from(leg14), c);
    int n = (int)(tail - c);
    for (int i = 0; i < n; i++) {
        out20[2 * i + 0] = c[i].x;
        out20[2 * i + 1] = c[i].y;
    }
    return n;
}
// one_leg.cu:167-208 in the femur plane (x already relative to the coxa joint)
int ref_eval_plane_reach(float x, float y, const float* leg14) {
    return eval_plane_circles<REACH_USECASE>(x, y, leg_from(leg14)) ? 1 : 0;
}
int ref_eval_plane_dist(float x, float y, const float* leg14, float* out2) {
    bool v = eval_plane_circles<DIST_USECASE>(x, y, leg_from(leg14));
    out2[0] = x;
    out2[1] = y;
    return v ? 1 : 0;
}

// unified_math_cuda.cu.h:13-83
void ref_qt_rotate(const float* q4, const float* v3, float* out3) {
    float3 r = qtRotate(quat_from(q4), make_float3(v3[0], v3[1], v3[2]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void ref_qt_invert(const float* q4, float* out4) {
    Quaternion r = qtInvert(quat_from(q4));
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
void ref_qt_multiply(const float* a4, const float* b4, float* out4) {
    Quaternion r = qtMultiply(quat_from(a4), quat_from(b4));
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
void ref_quat_from_vect_angle(const float* axis3, float angle, float* out4) {
    Quaternion r = quatFromVectAngle(make_float3(axis3[0], axis3[1], axis3[2]), angle);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
void ref_rpy_from_quat(const float* q4, float* out3) {
    float3 r = rpyFromQuat(quat_from(q4));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
// octree_util.cu.h:164-172
void ref_rpy_to_quat(float r, float p, float y, float* out4) {
    Quaternion q = RPYtoQuat(r, p, y);
    out4[0] = q.x; out4[1] = q.y; out4[2] = q.z; out4[3] = q.w;
}
// one_leg_global.cu:48-60
void ref_rotate_leg_data(const float* q4, const float* leg14, float* out14) {
    LegDimensions d = rotate_leg_data(quat_from(q4), leg_from(leg14));
    std::memcpy(out14, &d, sizeof(d));
}
// octree_util.cu.h:105-159; box = {cx,cy,cz, hx,hy,hz}
unsigned ref_create_child_box(const float* parent6, unsigned child, const uint8_t* small3,
                              float* child6, int* missing_quad) {
    Box p, c;
    std::memcpy(&p, parent6, sizeof(Box));
    bool small[3] = {small3[0] != 0, small3[1] != 0, small3[2] != 0};
    uchar missing = 0;
    unsigned r = CreateChildBox(p, c, 3, child, small, missing);
    std::memcpy(child6, &c, sizeof(Box));
    *missing_quad = missing;
    return r;
}
int ref_is_in_box(const float* v3, const float* box6) {
    Box b;
    std::memcpy(&b, box6, sizeof(Box));
    return isInBox(make_float3(v3[0], v3[1], v3[2]), b) ? 1 : 0;
}
int ref_sizeof_node() { return (int)sizeof(Node); }
int ref_sizeof_legdim() { return (int)sizeof(LegDimensions); }

} // extern "C"
