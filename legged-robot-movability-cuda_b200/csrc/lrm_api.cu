// lrm_api.cu — the extern "C" layer declared in include/lrm_c.h.
//
// Host-pointer calls follow the contract of the reference's apply_kernel
// (cross_compiled.cu:34-79): allocate device buffers, H2D, event-timed kernel, D2H, free — the
// returned kernel_ms is the kernel alone.  Device-pointer calls are asynchronous on the caller's
// stream.  There is no CPU fallback anywhere in this file: every compute entry point either runs
// the CUDA kernels or fails with LRM_ERR_CUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "kernels.h"
#include "leg_plan.h"
#include "lrm_c.h"

namespace {

thread_local std::string g_error;

int fail(lrm_status code, const std::string& msg) {
    g_error = msg;
    return (int)code;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_error = std::string(where) + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear the sticky-free error state
    return (int)LRM_ERR_CUDA;
}
#define LRM_CUDA(call, where)                            \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, where); \
    } while (0)

// Frees device scratch on every exit path of a host-pointer call.
struct DeviceScratch {
    void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    cudaError_t alloc(void** out, size_t bytes) {
        cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
        if (e == cudaSuccess) ptr[n++] = *out;
        return e;
    }
    ~DeviceScratch() {
        for (int i = 0; i < n; i++) cudaFree(ptr[i]);
    }
};

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    bool on = false;
    cudaError_t start(cudaStream_t s, bool want) {
        on = want;
        if (!on) return cudaSuccess;
        cudaError_t e = cudaEventCreate(&a);
        if (e != cudaSuccess) return e;
        e = cudaEventCreate(&b);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(a, s);
    }
    cudaError_t stop(cudaStream_t s, float* ms) {
        if (!on) return cudaSuccess;
        cudaError_t e = cudaEventRecord(b, s);
        if (e != cudaSuccess) return e;
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) return e;
        return cudaEventElapsedTime(ms, a, b);
    }
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};


// ---- pipelined staging for host-pointer calls --------------------------------------------------
// Chunk size: small enough that filling and draining the pipeline (one chunk up before the first
// kernel, one chunk down after the last) is a small share of a call, large enough that each copy
// runs at the link rate (>= 24 MiB on this box's PCIe 5 x16, tools/pcie_probe.py).
// lrm_set_option("staging_chunk_points") overrides it for measurements.
constexpr size_t kDefaultChunkPoints = size_t(1) << 21;  // 2 Mi points: 24 MiB up, 26 MiB down
constexpr int kSlots = 4;
std::atomic<size_t> g_chunk_points{kDefaultChunkPoints};
size_t chunk_points() { return g_chunk_points.load(std::memory_order_relaxed); }

struct Staging {
    size_t chunk_points = 0;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t uploaded[kSlots] = {}, computed[kSlots] = {}, drained[kSlots] = {};
    cudaEvent_t k0[kSlots] = {}, k1[kSlots] = {};
    float* d_in[kSlots] = {};
    float* d_vec[kSlots] = {};
    uint8_t* d_flag[kSlots] = {};
    cudaError_t init(size_t chunk, int nslots, bool want_vec, bool want_flag) {
        cudaError_t e;
        chunk_points = chunk;
        if ((e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int i = 0; i < nslots; i++) {
            if ((e = cudaEventCreateWithFlags(&uploaded[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreate(&k0[i])) != cudaSuccess) return e;
            if ((e = cudaEventCreate(&k1[i])) != cudaSuccess) return e;
            if ((e = cudaMalloc((void**)&d_in[i], chunk * 12)) != cudaSuccess) return e;
            if (want_vec && (e = cudaMalloc((void**)&d_vec[i], chunk * 12)) != cudaSuccess) return e;
            if (want_flag && (e = cudaMalloc((void**)&d_flag[i], chunk)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    ~Staging() {
        for (int i = 0; i < kSlots; i++) {
            if (d_in[i]) cudaFree(d_in[i]);
            if (d_vec[i]) cudaFree(d_vec[i]);
            if (d_flag[i]) cudaFree(d_flag[i]);
            if (uploaded[i]) cudaEventDestroy(uploaded[i]);
            if (computed[i]) cudaEventDestroy(computed[i]);
            if (drained[i]) cudaEventDestroy(drained[i]);
            if (k0[i]) cudaEventDestroy(k0[i]);
            if (k1[i]) cudaEventDestroy(k1[i]);
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_cmp) cudaStreamDestroy(s_cmp);
        if (s_out) cudaStreamDestroy(s_out);
    }
};

int staged_one_leg(int mode, const lrm::LegPlan& plan, const float* xyz, size_t n, float* out_xyz,
                   uint8_t* flags, cudaStream_t user_stream, float* kernel_ms) {
    const bool want_vec = (mode & lrm::kModeDist) != 0;
    const bool want_flag = flags != nullptr;
    const size_t kChunkPoints = chunk_points();
    const size_t chunk = n < kChunkPoints ? n : kChunkPoints;
    const size_t nchunks = (n + chunk - 1) / chunk;
    // Large sweeps reuse one set of streams / events / device buffers per device for the life of
    // the process (allocating ~600 MiB per call would cost as much as the copies); small calls
    // use a throw-away set sized to the call, like apply_kernel's malloc + free.
    static std::mutex pool_mutex;
    static Staging* pool[64] = {nullptr};
    std::unique_lock<std::mutex> lock(pool_mutex, std::defer_lock);
    Staging local;
    Staging* stp = &local;
    if (chunk == kChunkPoints) {
        lock.lock();
        int dev = 0;
        cudaGetDevice(&dev);
        Staging*& slot = pool[dev & 63];
        if (slot && slot->chunk_points != kChunkPoints) {  // the option changed: calls are synchronous, the set is idle
            delete slot;
            slot = nullptr;
        }
        if (!slot) {
            slot = new Staging();
            cudaError_t e = slot->init(kChunkPoints, kSlots, true, true);
            if (e != cudaSuccess) {
                delete slot;
                slot = nullptr;
                return cuda_fail(e, "staging buffers");
            }
        }
        stp = slot;
    } else {
        LRM_CUDA(local.init(chunk, nchunks < (size_t)kSlots ? (int)nchunks : kSlots, want_vec, want_flag),
                 "staging buffers");
    }
    Staging& st = *stp;
    // order after whatever the caller queued on its stream
    cudaEvent_t begin;
    LRM_CUDA(cudaEventCreateWithFlags(&begin, cudaEventDisableTiming), "cudaEventCreate");
    cudaEventRecord(begin, user_stream);
    cudaStreamWaitEvent(st.s_in, begin, 0);
    cudaStreamWaitEvent(st.s_cmp, begin, 0);
    cudaStreamWaitEvent(st.s_out, begin, 0);
    cudaEventDestroy(begin);

    double total_ms = 0.0;
    for (size_t c = 0; c < nchunks; c++) {
        const int slot = (int)(c % kSlots);
        const size_t first = c * chunk, cnt = (n - first < chunk) ? n - first : chunk;
        if (c >= (size_t)kSlots) {
            // the slot is free once its previous download has drained; collect its kernel time
            LRM_CUDA(cudaEventSynchronize(st.drained[slot]), "pipeline drain");
            float ms = 0.f;
            if (kernel_ms && cudaEventElapsedTime(&ms, st.k0[slot], st.k1[slot]) == cudaSuccess) total_ms += ms;
        }
        LRM_CUDA(cudaMemcpyAsync(st.d_in[slot], xyz + 3 * first, cnt * 12, cudaMemcpyHostToDevice, st.s_in),
                 "H2D points");
        cudaEventRecord(st.uploaded[slot], st.s_in);
        cudaStreamWaitEvent(st.s_cmp, st.uploaded[slot], 0);
        cudaEventRecord(st.k0[slot], st.s_cmp);
        LRM_CUDA(lrm::launch_one_leg_aos(mode, plan, st.d_in[slot], st.d_vec[slot], st.d_flag[slot], cnt,
                                         st.s_cmp, n),
                 "one-leg kernel launch");
        cudaEventRecord(st.k1[slot], st.s_cmp);
        cudaEventRecord(st.computed[slot], st.s_cmp);
        cudaStreamWaitEvent(st.s_out, st.computed[slot], 0);
        if (want_vec)
            LRM_CUDA(cudaMemcpyAsync(out_xyz + 3 * first, st.d_vec[slot], cnt * 12, cudaMemcpyDeviceToHost,
                                     st.s_out),
                     "D2H vectors");
        if (want_flag)
            LRM_CUDA(cudaMemcpyAsync(flags + first, st.d_flag[slot], cnt, cudaMemcpyDeviceToHost, st.s_out),
                     "D2H flags");
        cudaEventRecord(st.drained[slot], st.s_out);
        // (the next upload into this slot is issued only after the host has seen drained[slot],
        // which implies this chunk's kernel has finished with d_in[slot])
    }
    LRM_CUDA(cudaStreamSynchronize(st.s_out), "pipeline synchronize");
    LRM_CUDA(cudaStreamSynchronize(st.s_cmp), "pipeline synchronize");
    if (kernel_ms) {
        const size_t tail = nchunks < (size_t)kSlots ? nchunks : (size_t)kSlots;
        for (size_t k = 0; k < tail; k++) {
            const int slot = (int)((nchunks - 1 - k) % kSlots);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, st.k0[slot], st.k1[slot]) == cudaSuccess) total_ms += ms;
        }
        *kernel_ms = (float)total_ms;
    }
    return LRM_OK;
}

// Shared body of lrm_reach / lrm_dist / lrm_reach_dist.
int one_leg_call(int mode, const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat,
                 float* out_xyz, uint8_t* flags, int on_device, void* stream_v, float* kernel_ms) {
    if (!leg) return fail(LRM_ERR_INVALID, "leg is NULL");
    if (n && !xyz) return fail(LRM_ERR_INVALID, "xyz is NULL");
    if (n && (mode & lrm::kModeDist) && !out_xyz) return fail(LRM_ERR_INVALID, "out_xyz is NULL");
    if (n && mode == lrm::kModeReach && !flags) return fail(LRM_ERR_INVALID, "flags is NULL");
    if (n && mode == lrm::kModeBoth && !flags) return fail(LRM_ERR_INVALID, "reach_flags is NULL");
    if (kernel_ms) *kernel_ms = 0.f;

    lrm::LegPlan plan;
    lrm::build_leg_plan(*leg, quat, &plan);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

    if (on_device) {
        EventPair ev;
        LRM_CUDA(ev.start(stream, kernel_ms != nullptr), "cudaEventRecord");
        LRM_CUDA(lrm::launch_one_leg_aos(mode, plan, xyz, out_xyz, flags, n, stream),
                 "one-leg kernel launch");
        LRM_CUDA(ev.stop(stream, kernel_ms), "one-leg kernel");
        return LRM_OK;
    }

    // host pointers: apply_kernel's malloc / H2D / kernel / D2H / free (cross_compiled.cu:41-77),
    // pipelined: the array is cut into chunks so that the upload of chunk k+1, the kernel of chunk
    // k and the download of chunk k-1 overlap on three streams (full-duplex PCIe when the caller's
    // buffers are pinned; with pageable memory the copies serialise but stay correct).
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return cuda_fail(ce != cudaSuccess ? ce : cudaErrorNoDevice,
                         "no CUDA device (this library has no CPU path)");
    if (n == 0) return LRM_OK;
    return staged_one_leg(mode, plan, xyz, n, out_xyz, flags, stream, kernel_ms);
}

}  // namespace

extern "C" {

int lrm_abi_version(void) { return LRM_ABI_VERSION; }
const char* lrm_last_error(void) { return g_error.c_str(); }

size_t lrm_set_fast_path_min_points(size_t n) { return lrm::set_fast_path_min_points(n); }

int lrm_set_option(const char* name, double value, double* previous) {
    if (!name) return fail(LRM_ERR_INVALID, "name is NULL");
    const std::string k(name);
    double prev = 0.0;
    if (k == "fast_path_min_points") {
        if (!(value >= 0)) return fail(LRM_ERR_INVALID, "fast_path_min_points must be >= 0");
        prev = (double)lrm::set_fast_path_min_points((size_t)value);
    } else if (k == "sweep") {
        if (value != 0 && value != 1 && value != 2) return fail(LRM_ERR_INVALID, "sweep must be 0, 1 or 2");
        prev = lrm::set_sweep_mode((int)value);
    } else if (k == "tier_kernel") {
        if (value != 0 && value != 1) return fail(LRM_ERR_INVALID, "tier_kernel must be 0 or 1");
        prev = lrm::set_tier_kernel((int)value);
    } else if (k == "tier_chunk_shift") {
        if (!(value >= 0 && value <= 8)) return fail(LRM_ERR_INVALID, "tier_chunk_shift must be 0..8");
        prev = lrm::set_tier_chunk_shift((int)value);
    } else if (k == "volume_cell_mm" || k == "volume_dim") {
        float cell;
        int dim;
        lrm::get_choice_volume_shape(&cell, &dim);
        prev = k == "volume_dim" ? (double)dim : (double)cell;
        if (lrm::set_choice_volume_shape(k == "volume_dim" ? cell : (float)value, k == "volume_dim" ? (int)value : dim) != 0)
            return fail(LRM_ERR_INVALID, "volume_cell_mm must be 0.5..64, volume_dim a multiple of 4 in 16..1024");
    } else if (k == "volume_bricks") {
        prev = lrm::set_volume_bricks(value != 0 ? 1 : 0);
    } else if (k == "staging_chunk_points") {
        if (!(value >= 4096 && value <= (double)(size_t(1) << 30)))
            return fail(LRM_ERR_INVALID, "staging_chunk_points must be 4096..2^30");
        prev = (double)g_chunk_points.exchange((size_t)value & ~size_t(15));
    } else if (k == "skeleton") {
        const int r = lrm::set_skeleton(value != 0 ? 1 : 0);
        if (r < 0) return fail(LRM_ERR_UNSUPPORTED, "skeleton needs a measurement build (-DLRM_ENABLE_SKELETON)");
        prev = r;
    } else {
        return fail(LRM_ERR_INVALID, "unknown option: " + k);
    }
    if (previous) *previous = prev;
    return LRM_OK;
}

int lrm_get_stat(const char* name, double* value) {
    if (!name || !value) return fail(LRM_ERR_INVALID, "NULL argument");
    const std::string k(name);
    if (k == "table_builds") {
        *value = (double)lrm::table_builds();
    } else if (k == "volume_cell_mm" || k == "volume_dim") {
        float cell;
        int dim;
        lrm::get_choice_volume_shape(&cell, &dim);
        *value = k == "volume_dim" ? (double)dim : (double)cell;
    } else if (k == "volume_builds") {
        *value = (double)lrm::volume_builds_done();
    } else if (k == "volume_bricks" || k == "volume_brick_capacity") {
        unsigned used, cap;
        lrm::get_brick_stats(&used, &cap);
        *value = k == "volume_bricks" ? (double)used : (double)cap;
    } else {
        return fail(LRM_ERR_INVALID, "unknown statistic: " + k);
    }
    return LRM_OK;
}

int lrm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int lrm_set_device(int device) {
    LRM_CUDA(cudaSetDevice(device), "cudaSetDevice");
    return LRM_OK;
}

int lrm_default_leg(int robot, float azimuth, lrm_leg_t* out) {
    if (!out) return fail(LRM_ERR_INVALID, "out is NULL");
    if (robot != 0 && robot != 1) return fail(LRM_ERR_INVALID, "robot must be 0 (moonbot) or 1 (M2)");
    lrm::default_leg(robot, azimuth, out);
    return LRM_OK;
}

int lrm_reach(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, uint8_t* flags,
              int on_device, void* stream, float* kernel_ms) {
    return one_leg_call(lrm::kModeReach, xyz, n, leg, quat, nullptr, flags, on_device, stream,
                        kernel_ms);
}
int lrm_dist(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, float* out_xyz,
             uint8_t* flags, int on_device, void* stream, float* kernel_ms) {
    return one_leg_call(lrm::kModeDist, xyz, n, leg, quat, out_xyz, flags, on_device, stream,
                        kernel_ms);
}
int lrm_reach_dist(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat,
                   uint8_t* reach_flags, float* out_xyz, int on_device, void* stream,
                   float* kernel_ms) {
    return one_leg_call(lrm::kModeBoth, xyz, n, leg, quat, out_xyz, reach_flags, on_device, stream,
                        kernel_ms);
}

int lrm_reach_dist_soa(const float* x, const float* y, const float* z, size_t n,
                       const lrm_leg_t* leg, const float* quat, uint8_t* reach_flags, float* dx,
                       float* dy, float* dz, void* stream_v, float* kernel_ms) {
    if (!leg) return fail(LRM_ERR_INVALID, "leg is NULL");
    if (n && (!x || !y || !z)) return fail(LRM_ERR_INVALID, "input plane is NULL");
    const bool want_vec = dx || dy || dz;
    if (want_vec && !(dx && dy && dz))
        return fail(LRM_ERR_INVALID, "dx, dy, dz must be all set or all NULL");
    if (n && !want_vec && !reach_flags) return fail(LRM_ERR_INVALID, "no output requested");
    if (kernel_ms) *kernel_ms = 0.f;
    lrm::LegPlan plan;
    lrm::build_leg_plan(*leg, quat, &plan);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    EventPair ev;
    LRM_CUDA(ev.start(stream, kernel_ms != nullptr), "cudaEventRecord");
    LRM_CUDA(lrm::launch_one_leg_soa(plan, x, y, z, dx, dy, dz, reach_flags, n, stream),
             "one-leg SoA kernel launch (planes must be 16-byte aligned)");
    LRM_CUDA(ev.stop(stream, kernel_ms), "one-leg SoA kernel");
    return LRM_OK;
}

int lrm_forward_kine(const float* angles, size_t n, const lrm_leg_t* leg, float* out_xyz,
                     int on_device, void* stream_v, float* kernel_ms) {
    if (!leg) return fail(LRM_ERR_INVALID, "leg is NULL");
    if (n && (!angles || !out_xyz)) return fail(LRM_ERR_INVALID, "NULL buffer");
    if (kernel_ms) *kernel_ms = 0.f;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    DeviceScratch scratch;
    const float* d_in = angles;
    float* d_out = out_xyz;
    if (!on_device) {
        float* tmp = nullptr;
        LRM_CUDA(scratch.alloc((void**)&tmp, n * 12), "cudaMalloc angles");
        LRM_CUDA(scratch.alloc((void**)&d_out, n * 12), "cudaMalloc xyz");
        LRM_CUDA(cudaMemcpyAsync(tmp, angles, n * 12, cudaMemcpyHostToDevice, stream), "H2D angles");
        d_in = tmp;
    }
    {
        EventPair ev;
        LRM_CUDA(ev.start(stream, kernel_ms != nullptr), "cudaEventRecord");
        LRM_CUDA(lrm::launch_forward_kine(d_in, *leg, d_out, n, stream), "forward-kine launch");
        LRM_CUDA(ev.stop(stream, kernel_ms), "forward-kine kernel");
    }
    if (!on_device) {
        LRM_CUDA(cudaMemcpyAsync(out_xyz, d_out, n * 12, cudaMemcpyDeviceToHost, stream), "D2H xyz");
        LRM_CUDA(cudaStreamSynchronize(stream), "stream synchronize");
    }
    return LRM_OK;
}

int lrm_recurs(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, int max_depth,
               float* out_xyz, int on_device, void* stream_v, float* kernel_ms) {
    if (!leg) return fail(LRM_ERR_INVALID, "leg is NULL");
    if (n && (!xyz || !out_xyz)) return fail(LRM_ERR_INVALID, "NULL buffer");
    if (max_depth < 0 || max_depth > 32) return fail(LRM_ERR_INVALID, "max_depth must be 0..32");
    if (kernel_ms) *kernel_ms = 0.f;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    lrm::LegPlan plan;
    lrm::build_leg_plan(*leg, quat, &plan);
    DeviceScratch scratch;
    const float* d_in = xyz;
    float* d_out = out_xyz;
    if (!on_device) {
        int count = 0;
        cudaError_t ce = cudaGetDeviceCount(&count);
        if (ce != cudaSuccess || count == 0)
            return cuda_fail(ce != cudaSuccess ? ce : cudaErrorNoDevice,
                             "no CUDA device (this library has no CPU path)");
        float* tmp = nullptr;
        LRM_CUDA(scratch.alloc((void**)&tmp, n * 12), "cudaMalloc points");
        LRM_CUDA(scratch.alloc((void**)&d_out, n * 12), "cudaMalloc output");
        LRM_CUDA(cudaMemcpyAsync(tmp, xyz, n * 12, cudaMemcpyHostToDevice, stream), "H2D points");
        // untouched entries keep the caller's values
        LRM_CUDA(cudaMemcpyAsync(d_out, out_xyz, n * 12, cudaMemcpyHostToDevice, stream), "H2D output");
        d_in = tmp;
    }
    {
        EventPair ev;
        LRM_CUDA(ev.start(stream, kernel_ms != nullptr), "cudaEventRecord");
        LRM_CUDA(lrm::launch_recurs(plan, d_in, d_out, n, max_depth, stream), "recurs kernel launch");
        LRM_CUDA(ev.stop(stream, kernel_ms), "recurs kernel");
    }
    if (!on_device) {
        LRM_CUDA(cudaMemcpyAsync(out_xyz, d_out, n * 12, cudaMemcpyDeviceToHost, stream), "D2H output");
        LRM_CUDA(cudaStreamSynchronize(stream), "stream synchronize");
    }
    return LRM_OK;
}

int lrm_make_lattice(float* out_xyz, const float lo[3], const float step[3], const uint32_t dims[3],
                     size_t first, size_t count, void* stream) {
    if (!lo || !step || !dims) return fail(LRM_ERR_INVALID, "NULL lattice description");
    if (count && !out_xyz) return fail(LRM_ERR_INVALID, "out_xyz is NULL");
    const size_t total = (size_t)dims[0] * dims[1] * dims[2];
    if (dims[0] == 0 || dims[1] == 0 || dims[2] == 0 || first + count > total)
        return fail(LRM_ERR_INVALID, "lattice range out of bounds");
    LRM_CUDA(lrm::launch_lattice(out_xyz, lo, step, dims, first, count,
                                 static_cast<cudaStream_t>(stream)),
             "lattice kernel launch");
    return LRM_OK;
}

int lrm_full_struct_orientations(float* out, int capacity) {
    // several_leg.cu:811-857
    if (!out || capacity < 45) return fail(LRM_ERR_INVALID, "need room for 45 quaternions");
    const float pi = 3.14159265358979323846264338327950288419716939937510582097f;
    const float ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
    float q_init[4], tmp[4], q_roll[4], q_pitch[4], q_yaw[4];
    lrm::quat_from_vect_angle(az, 0.f, q_init);
    const float r_min = -pi / 8, r_max = pi / 8, p_min = -pi / 8, p_max = +pi / 8;
    const float y_min = 0, y_max = pi / 2;
    int k = 0;
    for (int i = 0; i <= 2; i++) {
        const float roll = r_min + (r_max - r_min) * ((float)i / 2.f);
        lrm::quat_from_vect_angle(ax, roll, tmp);
        lrm::quat_multiply(tmp, q_init, q_roll);
        for (int j = 0; j <= 2; j++) {
            const float pitch = p_min + (p_max - p_min) * ((float)j / 2.f);
            lrm::quat_from_vect_angle(ay, pitch, tmp);
            lrm::quat_multiply(tmp, q_roll, q_pitch);
            for (int m = 0; m <= 4; m++) {
                const float yaw = y_min + (y_max - y_min) * ((float)m / 4.f);
                lrm::quat_from_vect_angle(az, yaw, tmp);
                lrm::quat_multiply(tmp, q_pitch, q_yaw);
                std::memcpy(out + 4 * k, q_yaw, sizeof q_yaw);
                k++;
            }
        }
    }
    return LRM_OK;
}

int lrm_rpy_to_quat(float roll, float pitch, float yaw, float out[4]) {
    if (!out) return fail(LRM_ERR_INVALID, "out is NULL");
    const float ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
    float q_roll[4], q_pitch[4], q_yaw[4], tmp[4];
    lrm::quat_from_vect_angle(ax, roll, q_roll);
    lrm::quat_from_vect_angle(ay, pitch, tmp);
    lrm::quat_multiply(tmp, q_roll, q_pitch);
    lrm::quat_from_vect_angle(az, yaw, tmp);
    lrm::quat_multiply(tmp, q_pitch, q_yaw);
    std::memcpy(out, q_yaw, sizeof q_yaw);
    return LRM_OK;
}

static int positionability_call(const float* bodies, size_t nb, const float* map, size_t nt,
                                const lrm_leg_t* legs, int nlegs, const float* quats, int nq,
                                const lrm_posit_opts_t* opts, uint8_t* standable, int on_device,
                                void* stream_v, float* kernel_ms, double* stats) {
    if (!legs || nlegs <= 0 || nlegs > 8) return fail(LRM_ERR_INVALID, "1..8 legs required");
    if (!quats || nq <= 0 || nq > 254) return fail(LRM_ERR_INVALID, "1..254 orientations required");
    if (nb && (!bodies || !standable)) return fail(LRM_ERR_INVALID, "NULL body buffer");
    if (nt && !map) return fail(LRM_ERR_INVALID, "map is NULL");
    if (opts && opts->first_hit_only) return fail(LRM_ERR_INVALID, "first_hit_only must be 0");
    if (kernel_ms) *kernel_ms = 0.f;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

    lrm::PositParams p;
    p.nb = nb, p.nt = nt, p.legs = legs, p.nlegs = nlegs, p.quats = quats, p.nq = nq;
    p.pre_cull = opts ? opts->pre_cull : 0;
    p.stats = stats;
    DeviceScratch scratch;
    if (on_device) {
        p.bodies = bodies, p.map = map, p.standable = standable;
    } else {
        int count = 0;
        cudaError_t ce = cudaGetDeviceCount(&count);
        if (ce != cudaSuccess || count == 0)
            return cuda_fail(ce != cudaSuccess ? ce : cudaErrorNoDevice,
                             "no CUDA device (this library has no CPU path)");
        float *d_b = nullptr, *d_m = nullptr;
        uint8_t* d_s = nullptr;
        LRM_CUDA(scratch.alloc((void**)&d_b, nb * 12), "cudaMalloc bodies");
        LRM_CUDA(scratch.alloc((void**)&d_m, nt * 12), "cudaMalloc map");
        LRM_CUDA(scratch.alloc((void**)&d_s, nb), "cudaMalloc result");
        LRM_CUDA(cudaMemcpyAsync(d_b, bodies, nb * 12, cudaMemcpyHostToDevice, stream), "H2D bodies");
        LRM_CUDA(cudaMemcpyAsync(d_m, map, nt * 12, cudaMemcpyHostToDevice, stream), "H2D map");
        p.bodies = d_b, p.map = d_m, p.standable = d_s;
    }
    LRM_CUDA(lrm::run_positionability(p, stream, kernel_ms), "positionability");
    if (!on_device) {
        LRM_CUDA(cudaMemcpyAsync(standable, p.standable, nb, cudaMemcpyDeviceToHost, stream),
                 "D2H result");
        LRM_CUDA(cudaStreamSynchronize(stream), "stream synchronize");
    }
    return LRM_OK;
}

int lrm_positionability(const float* bodies, size_t nb, const float* map, size_t nt,
                        const lrm_leg_t* legs, int nlegs, const float* quats, int nq,
                        const lrm_posit_opts_t* opts, uint8_t* standable, int on_device,
                        void* stream_v, float* kernel_ms) {
    return positionability_call(bodies, nb, map, nt, legs, nlegs, quats, nq, opts, standable, on_device, stream_v,
                                kernel_ms, nullptr);
}

int lrm_positionability_counts(const float* bodies, size_t nb, const float* map, size_t nt,
                               const lrm_leg_t* legs, int nlegs, const float* quats, int nq,
                               const lrm_posit_opts_t* opts, uint8_t* standable, double counts[3],
                               int on_device, void* stream_v) {
    if (!counts) return fail(LRM_ERR_INVALID, "counts is NULL");
    counts[0] = counts[1] = counts[2] = 0.0;
    return positionability_call(bodies, nb, map, nt, legs, nlegs, quats, nq, opts, standable, on_device, stream_v,
                                nullptr, counts);
}

int lrm_oct_sharded(const float* footholds, size_t nt, const lrm_leg_t* leg, int max_depth, int shard, int nshards,
                    float* out_xyz, size_t cap, size_t* count, size_t child_counts[8], int on_device,
                    void* stream_v, float* kernel_ms) {
    if (!leg || !count) return fail(LRM_ERR_INVALID, "leg / count is NULL");
    if (nt && !footholds) return fail(LRM_ERR_INVALID, "footholds is NULL");
    if (cap && !out_xyz) return fail(LRM_ERR_INVALID, "out_xyz is NULL");
    if (max_depth < 0 || max_depth > 16) return fail(LRM_ERR_INVALID, "max_depth must be 0..16");
    if (nshards < 1 || shard < 0 || shard >= nshards) return fail(LRM_ERR_INVALID, "shard must be in [0, nshards)");
    *count = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return cuda_fail(ce != cudaSuccess ? ce : cudaErrorNoDevice,
                         "no CUDA device (this library has no CPU path)");
    DeviceScratch scratch;
    const float* d_foot = footholds;
    if (!on_device) {
        float* tmp = nullptr;
        LRM_CUDA(scratch.alloc((void**)&tmp, nt * 12), "cudaMalloc footholds");
        LRM_CUDA(cudaMemcpyAsync(tmp, footholds, nt * 12, cudaMemcpyHostToDevice, stream), "H2D footholds");
        d_foot = tmp;
    }
    std::vector<float> centres;
    LRM_CUDA(lrm::run_octree(d_foot, nt, *leg, max_depth, &centres, stream, kernel_ms, shard, nshards, child_counts),
             "octree");
    *count = centres.size() / 3;
    const size_t n = *count < cap ? *count : cap;
    if (n) std::memcpy(out_xyz, centres.data(), n * 12);
    return LRM_OK;
}

int lrm_oct(const float* footholds, size_t nt, const lrm_leg_t* leg, int max_depth, float* out_xyz,
            size_t cap, size_t* count, int on_device, void* stream_v, float* kernel_ms) {
    return lrm_oct_sharded(footholds, nt, leg, max_depth, 0, 1, out_xyz, cap, count, nullptr, on_device, stream_v,
                           kernel_ms);
}

int lrm_oct_children(const float* footholds, size_t nt, const lrm_leg_t* leg, const float parent_box6[6],
                     int parent_validity, uint8_t* out_flags, float* out_boxes, int on_device, void* stream_v) {
    if (!leg || !parent_box6 || !out_flags || !out_boxes) return fail(LRM_ERR_INVALID, "NULL argument");
    if (nt && !footholds) return fail(LRM_ERR_INVALID, "footholds is NULL");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return cuda_fail(ce != cudaSuccess ? ce : cudaErrorNoDevice,
                         "no CUDA device (this library has no CPU path)");
    DeviceScratch scratch;
    const float* d_foot = footholds;
    if (!on_device) {
        float* tmp = nullptr;
        LRM_CUDA(scratch.alloc((void**)&tmp, nt * 12), "cudaMalloc footholds");
        LRM_CUDA(cudaMemcpyAsync(tmp, footholds, nt * 12, cudaMemcpyHostToDevice, stream), "H2D footholds");
        d_foot = tmp;
    }
    LRM_CUDA(lrm::run_octree_children(d_foot, nt, *leg, parent_box6, parent_validity, out_flags, out_boxes, stream),
             "octree children");
    return LRM_OK;
}

}  // extern "C"
