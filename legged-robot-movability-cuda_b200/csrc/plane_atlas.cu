// plane_atlas.cu — builds and caches the plane atlas of a LegPlan (see leg_math.cuh).
//
// One byte per cell of a regular grid over the femur plane: the certified outcome of plane_clamp
// (valid bit, sector, winning candidate) or "impure".  Built on the device by probing every cell
// centre with the instrumented evaluation (plane_probe): a cell is certified only when the smallest
// decision margin exceeds the cell's half diagonal plus a float-rounding allowance (kAtlasNeed*).  The atlas is a
// pure accelerator: the streaming kernel falls back to the full evaluation for every point it
// cannot certify.  16 MiB per (leg, orientation); a small per-device LRU keeps the last few.
#include <cuda_runtime.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.h"
#include "leg_math.cuh"

namespace lrm {

namespace {

constexpr int kAtlasDim = 4096;       // cells per side
constexpr float kAtlasCell = 0.5f;    // mm
constexpr float kAtlasOrigin = -0.5f * kAtlasDim * kAtlasCell;  // [-1024, 1024) mm in X and Y

__global__ void atlas_build_kernel(const __grid_constant__ LegPlan L, unsigned char* __restrict__ blocked,
                                   unsigned char* __restrict__ linear, int dim, float origin, float cell) {
    __shared__ SectorTable table;
    fill_sector_table(L, &table, threadIdx.x, blockDim.x);
    __syncthreads();
    const size_t total = (size_t)dim * dim;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // half diagonal + allowance for float rounding of the margins and of the cell index (the
    // texture unit resolves cell coordinates to 1/256 of a cell)
    const float need = kAtlasNeedFactor * cell + kAtlasNeedSlack;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ix = (int)(i % dim), iy = (int)(i / dim);
        const float X = origin + ((float)ix + 0.5f) * cell, Y = origin + ((float)iy + 0.5f) * cell;
        const PlaneProbe pr = plane_probe(L, table, X, Y);
        const unsigned char v = (unsigned char)atlas_cell_byte(pr, need);
        blocked[atlas_index(dim, ix, iy)] = v;
        linear[i] = v;
    }
}

// Coarse pass of the volume build: one thread per block of 4 x 4 x 4 cubes.  The certification is
// scale-free: a block whose centre decides both the choice and the reach bits for the WHOLE block
// (no refinement needed) hands that byte to its 64 cubes; the fine pass below only looks at the
// cubes of the other blocks — the shells around the decision surfaces.
__global__ void __launch_bounds__(128)
    volume_coarse_kernel(const __grid_constant__ LegPlan L, const __grid_constant__ FastTables FT,
                         unsigned* __restrict__ linear, unsigned char* __restrict__ block_done, int dim,
                         float cell) {
    __shared__ SectorTable table;
    fill_sector_table(L, &table, threadIdx.x, blockDim.x);
    __syncthreads();
    const int bd = dim >> 2;
    const size_t total = (size_t)bd * bd * bd;
    const float half = 0.5f * (float)dim;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int bx = (int)(i % bd), by = (int)((i / bd) % bd), bz = (int)(i / ((size_t)bd * bd));
        const float x0 = ((float)(4 * bx) - half) * cell, y0 = ((float)(4 * by) - half - kVolShiftY) * cell,
                    z0 = ((float)(4 * bz) - half) * cell;
        const unsigned w = coarse_block_word(L, table, FT, x0, y0, z0, 4.f * cell);
        block_done[i] = w != 0u ? 1 : 0;
        if (w != 0u) {
            const uint4 four = make_uint4(w, w, w, w);  // four texels along x (dim % 4 == 0)
            for (int dz = 0; dz < 4; dz++)
                for (int dy = 0; dy < 4; dy++)
                    *reinterpret_cast<uint4*>(
                        linear + ((size_t)(4 * bz + dz) * dim + (size_t)(4 * by + dy)) * dim + 4 * bx) = four;
        }
    }
}

// Classic texel of the cube at (x0, y0, z0), side `cell`; ALL 32 lanes of a warp call it together,
// one cube per lane (live = false: no cube).  Cubes whose centre cannot decide are refined on 4^3
// sub-cubes; those cubes hug the decision surfaces (a few lanes per warp), so the warp refines them
// one after the other with all 32 lanes, two sub-cubes each, instead of leaving 29 lanes idle while
// three of them run 64 probes.  pad_h: see choice_cell_first.
__device__ __forceinline__ unsigned cube_word_coop(const LegPlan& L, const SectorTable& table, const FastTables& FT,
                                                   const AtlasView& atlas, float x0, float y0, float z0, float cell,
                                                   float pad_h, bool live, int lane) {
    static_assert(kVolSub * kVolSub * kVolSub == 64, "two sub-cubes per lane");
    CellFirst f;
    f.byte = 0u, f.refine = false, f.direct = true, f.reach = 0u, f.reach_refine = false, f.reach_flip = false;
    if (live) f = choice_cell_first(L, table, FT, x0, y0, z0, cell, pad_h);
    unsigned need = __ballot_sync(0xffffffffu, f.refine);
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const float sx = __shfl_sync(0xffffffffu, x0, src), sy = __shfl_sync(0xffffffffu, y0, src),
                    sz = __shfl_sync(0xffffffffu, z0, src);
        const bool sdir = __shfl_sync(0xffffffffu, (int)f.direct, src) != 0;
        const bool ok = choice_cell_sub(L, table, sx, sy, sz, cell, lane, sdir, pad_h) &&
                        choice_cell_sub(L, table, sx, sy, sz, cell, lane + 32, sdir, pad_h);
        const bool all = __all_sync(0xffffffffu, ok);
        if (lane == src && !all) f.byte = 0u;
    }
    need = __ballot_sync(0xffffffffu, f.reach_refine);
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const float sx = __shfl_sync(0xffffffffu, x0, src), sy = __shfl_sync(0xffffffffu, y0, src),
                    sz = __shfl_sync(0xffffffffu, z0, src);
        const unsigned sbits = __shfl_sync(0xffffffffu, f.reach | (f.reach_flip ? 1u : 0u), src);
        const bool flip = (sbits & 1u) != 0u, valid = (sbits & kVolReachValue) != 0u;
        const bool ok = reach_cell_sub(L, table, sx, sy, sz, cell, lane, flip, valid, pad_h) &&
                        reach_cell_sub(L, table, sx, sy, sz, cell, lane + 32, flip, valid, pad_h);
        const bool all = __all_sync(0xffffffffu, ok);
        if (lane == src && !all) f.reach = 0u;
    }
    if (!live) return 0u;
    // plane label of the chosen solution (tier 0 of the sweep): the atlas cells under the cube's
    // plane rectangle all carry one certified label
    unsigned hi = 0u;
    if (f.byte & kVolPure) {
        float side;
        const CoxaPoint c = cube_centre(x0, y0, z0, cell, &side, pad_h);
        hi = cube_plane_scan(atlas, solution_plane_x(L, c, (f.byte & 1u) != 0u), c.z, side);
    }
    return f.byte | f.reach | (hi << 8);
}

// One lane per cube of the choice volume (x fastest, like the 3-D array upload).
__global__ void __launch_bounds__(128)
    volume_build_kernel(const __grid_constant__ LegPlan L, const __grid_constant__ FastTables FT, const AtlasView atlas,
                        unsigned* __restrict__ linear, const unsigned char* __restrict__ block_done, int dim,
                        float cell) {
    __shared__ SectorTable table;
    fill_sector_table(L, &table, threadIdx.x, blockDim.x);
    __syncthreads();
    const size_t total = (size_t)dim * dim * dim;
    const size_t stride = (size_t)gridDim.x * blockDim.x;  // a multiple of 32: whole warps step together
    const float half = 0.5f * (float)dim;
    const int lane = threadIdx.x & 31;
    const size_t rounds = (total + stride - 1) / stride;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t r = 0; r < rounds; r++, i += stride) {
        const int ix = (int)(i % dim), iy = (int)((i / dim) % dim), iz = (int)(i / ((size_t)dim * dim));
        // cubes of a block the coarse pass has settled are already written
        const bool live = i < total && !(block_done != nullptr &&
                                         block_done[((size_t)(iz >> 2) * (dim >> 2) + (iy >> 2)) * (dim >> 2) + (ix >> 2)]);
        const float x0 = ((float)ix - half) * cell, y0 = ((float)iy - half - kVolShiftY) * cell, z0 = ((float)iz - half) * cell;
        const unsigned w = cube_word_coop(L, table, FT, atlas, x0, y0, z0, cell, 0.f, live, lane);
        if (live) linear[i] = w;
    }
}

// Bricks (see leg_math.cuh): a warp reads 32 consecutive coarse texels, and for every cube among them
// that wants a brick computes its 64 fine texels — two per lane — with the same functions, the box
// widened by the coarse cube's pad.  A brick is kept if at least kBrickMinUseful of its fine cubes
// end up settled for tier 0 (choice and plane label); the coarse texel then becomes the pointer.
// Bricks are numbered by an atomic counter; cubes that come after the pool is full stay as they are.
constexpr int kBrickMinUseful = 8;
__global__ void __launch_bounds__(128)
    brick_build_kernel(const __grid_constant__ LegPlan L, const __grid_constant__ FastTables FT, const AtlasView atlas,
                       unsigned* __restrict__ linear, unsigned short* __restrict__ bricks, unsigned* __restrict__ counter,
                       unsigned capacity, int dim, float cell) {
    __shared__ SectorTable table;
    fill_sector_table(L, &table, threadIdx.x, blockDim.x);
    __syncthreads();
    const size_t total = (size_t)dim * dim * dim;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float half = 0.5f * (float)dim;
    const int lane = threadIdx.x & 31;
    const float hf = cell * (1.f / kBrickSub);
    for (size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~size_t(31); base < total; base += stride) {
        const size_t i = base + lane;
        const unsigned w = i < total ? linear[i] : 0u;
        unsigned need = __ballot_sync(0xffffffffu, i < total && brick_candidate(w));
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const size_t ci = base + src;
            const unsigned cw = __shfl_sync(0xffffffffu, w, src);
            const int ix = (int)(ci % dim), iy = (int)((ci / dim) % dim), iz = (int)(ci / ((size_t)dim * dim));
            const float x0 = ((float)ix - half) * cell, y0 = ((float)iy - half - kVolShiftY) * cell, z0 = ((float)iz - half) * cell;
            unsigned fw[2];
#pragma unroll 1
            for (int hlf = 0; hlf < 2; hlf++) {
                const unsigned slot = (unsigned)lane + 32u * hlf;
                const float xf = x0 + (float)(slot & 3u) * hf, yf = y0 + (float)((slot >> 2) & 3u) * hf,
                            zf = z0 + (float)(slot >> 4) * hf;
                if (cw & kVolPure) {  // warp-uniform: the choice holds for the whole coarse cube, only the label is missing
                    float side;
                    const CoxaPoint c = cube_centre(xf, yf, zf, hf, &side, cell);
                    const unsigned hi = cube_plane_scan(atlas, solution_plane_x(L, c, (cw & 1u) != 0u), c.z, side);
                    fw[hlf] = (cw & 0xffu) | (hi << 8);
                } else {
                    fw[hlf] = cube_word_coop(L, table, FT, atlas, xf, yf, zf, hf, cell, true, lane);
                }
            }
            const int useful = __popc(__ballot_sync(0xffffffffu, !brick_candidate(fw[0]))) +
                               __popc(__ballot_sync(0xffffffffu, !brick_candidate(fw[1])));
            if (useful >= kBrickMinUseful) {
                unsigned idx = 0u;
                if (lane == 0) idx = atomicAdd(counter, 1u);
                idx = __shfl_sync(0xffffffffu, idx, 0);
                if (idx < capacity) {
                    bricks[(size_t)idx * 64 + lane] = (unsigned short)fw[0];
                    bricks[(size_t)idx * 64 + 32 + lane] = (unsigned short)fw[1];
                    if (lane == 0) linear[ci] = brick_texel(idx, ix, iy, iz, cw);
                }
            }
        }
    }
}

// ---- cache ---------------------------------------------------------------------------------------
// Per device, a small LRU of (plan -> atlas, yaw tables, choice volume).  Nothing here makes a
// device-pointer call synchronous once a plan's tables exist, and a rebuild is ordered by events:
//   * a caller LEASES an entry (acquire_tables): the entry is pinned until release_tables, which
//     records an event on the caller's stream after its launches — eviction never touches a pinned
//     entry, so a second host thread cannot rebuild the tables between a lookup and the launch
//     that uses them;
//   * an entry's buffers are rewritten (eviction) only after the rebuilding stream has been made to
//     wait for every recorded use (cudaStreamWaitEvent; a device-wide wait only if more streams than
//     kUseSlots used the entry);
//   * a freshly built atlas is published by a `ready` event that other streams wait for on the
//     device; the host does not wait.
constexpr int kUseSlots = 4;
struct Entry {
    bool used = false;
    int device = -1;
    int pins = 0;
    unsigned long long stamp = 0;
    LegPlan plan;
    unsigned char* cells = nullptr;   // 8 x 4 blocked
    unsigned char* linear = nullptr;  // row-major staging copy for the texture upload
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
    FastTables tables;  // yaw-sector table of the same plan (host-built, ~0.5 ms: cached with the atlas)
    cudaEvent_t ready = nullptr;      // atlas build complete
    cudaStream_t ready_stream = nullptr;
    cudaStream_t use_stream[kUseSlots] = {};
    cudaEvent_t use_done[kUseSlots] = {};
    bool use_valid[kUseSlots] = {};
    bool use_overflow = false;        // more distinct streams than slots since the last rebuild
    // choice volume (built on first request, see get_choice_volume)
    cudaArray_t vol_array = nullptr;
    cudaTextureObject_t vol_tex = 0;
    int vol_dim = 0;
    float vol_cell = 0.f;
    bool vol_ready = false;
    // a build in flight on the entry's own stream (see get_choice_volume)
    bool vol_building = false;
    cudaStream_t vol_stream = nullptr;
    cudaEvent_t vol_done = nullptr;
    unsigned* vol_linear = nullptr;        // staging copy (32-bit texels), freed once the build has finished
    // bricks: fine texels under the cubes the coarse grid cannot settle (kept as long as the array)
    unsigned short* vol_bricks = nullptr;
    unsigned* vol_brick_count = nullptr;   // device word: bricks handed out by the last build (may exceed the pool)
    unsigned vol_brick_cap = 0;
    int vol_bricks_on = -1;                // the "volume_bricks" option the volume was built with
    float vol_build_ms = 0.f;              // wall time of the last build (host clock around the wait), diagnostics
};
// a build in flight must finish before its buffers or its plan's atlas go away
void settle_volume(Entry* c) {
    if (c->vol_building) {
        cudaEventSynchronize(c->vol_done);
        c->vol_building = false;
    }
    if (c->vol_linear) cudaFreeAsync(c->vol_linear, c->vol_stream);
    c->vol_linear = nullptr;
}
void release_volume(Entry* c) {
    settle_volume(c);
    if (c->vol_tex) cudaDestroyTextureObject(c->vol_tex);
    if (c->vol_array) cudaFreeArray(c->vol_array);
    if (c->vol_bricks) cudaFree(c->vol_bricks);
    if (c->vol_brick_count) cudaFree(c->vol_brick_count);
    c->vol_bricks = nullptr, c->vol_brick_count = nullptr, c->vol_brick_cap = 0;
    c->vol_tex = 0, c->vol_array = nullptr, c->vol_ready = false, c->vol_dim = 0;
}
// every stream that used the entry must have finished with it before its buffers are rewritten:
// ordered on the device where the uses are known, device-wide otherwise
cudaError_t order_after_uses(Entry* c, cudaStream_t stream) {
    if (c->use_overflow) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return e;
    }
    for (int k = 0; k < kUseSlots; k++)
        if (c->use_valid[k]) {
            cudaError_t e = cudaStreamWaitEvent(stream, c->use_done[k], 0);
            if (e != cudaSuccess) return e;
            c->use_valid[k] = false;
        }
    c->use_overflow = false;
    return cudaSuccess;
}
cudaError_t allocate(Entry* c) {
    const size_t bytes = (size_t)kAtlasDim * kAtlasDim;
    cudaError_t e = cudaMalloc((void**)&c->cells, bytes);
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void**)&c->linear, bytes);
    if (e != cudaSuccess) return e;
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    e = cudaMallocArray(&c->array, &fmt, kAtlasDim, kAtlasDim);
    if (e != cudaSuccess) return e;
    cudaResourceDesc res;
    std::memset(&res, 0, sizeof res);
    res.resType = cudaResourceTypeArray;
    res.res.array.array = c->array;
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;  // off the atlas -> 0 = impure
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    e = cudaCreateTextureObject(&c->tex, &res, &td, nullptr);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    for (int k = 0; k < kUseSlots; k++) {
        e = cudaEventCreateWithFlags(&c->use_done[k], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
void release(Entry* c) {
    release_volume(c);
    if (c->vol_stream) cudaStreamDestroy(c->vol_stream);
    if (c->vol_done) cudaEventDestroy(c->vol_done);
    c->vol_stream = nullptr, c->vol_done = nullptr;
    if (c->tex) cudaDestroyTextureObject(c->tex);
    if (c->array) cudaFreeArray(c->array);
    if (c->cells) cudaFree(c->cells);
    if (c->linear) cudaFree(c->linear);
    if (c->ready) cudaEventDestroy(c->ready);
    for (int k = 0; k < kUseSlots; k++) {
        if (c->use_done[k]) cudaEventDestroy(c->use_done[k]);
        c->use_done[k] = nullptr, c->use_valid[k] = false;
    }
    c->tex = 0, c->array = nullptr, c->cells = nullptr, c->linear = nullptr, c->ready = nullptr;
}

// Entries never move (leases hold pointers) and belong to ONE device for life: kCacheEntries per
// device (a hexapod sweeping its six legs in turn keeps all six), plus spill-over entries when
// every entry of a device is pinned at once.
constexpr int kCacheEntries = 8;
constexpr int kMaxDevices = 64;
std::vector<Entry*> g_cache[kMaxDevices];
unsigned long long g_clock = 0;
std::mutex g_mutex;
// The volume is built in the BACKGROUND while sweeps of the same plan keep running (two-tier sweep,
// same bits): a build that fills every SM (16 CTAs of 128 threads each) finishes in 40 ms but
// slows the foreground sweeps of those 40 ms by 50 x (measured: 0.5 ms -> 28-87 ms per call); with
// a few CTAs per SM the build takes longer and the foreground keeps most of the registers.
#ifndef LRM_VOL_BUILD_CTAS
#define LRM_VOL_BUILD_CTAS 3
#endif
constexpr int kVolBuildCtasPerSm = LRM_VOL_BUILD_CTAS;
std::atomic<int> g_vol_dim{512};
std::atomic<float> g_vol_cell{3.0f};
// bricks under the uncertified cubes (0: coarse grid only).  Off by default: on the bench lattice they
// halve the share of parked points (11.5 % -> 5.1 %) but the pointer chase costs what the smaller
// rings save (128.1 -> 128.9 Gpoints/s), for 0.25 s more background build and 1.4 GB per plan.
std::atomic<int> g_vol_bricks{0};
std::atomic<unsigned> g_last_bricks{0}, g_last_brick_cap{0};  // diagnostics of the last finished build
std::atomic<unsigned long long> g_vol_builds_done{0};         // volume builds seen finished since the library was loaded
std::atomic<unsigned long long> g_builds{0};  // atlas builds since the library was loaded (diagnostics / tests)

void fill_view(const Entry* hit, AtlasView* view) {
    view->cells = hit->cells;
    view->tex = hit->tex;
    view->inv_cell = 1.0f / kAtlasCell;
    view->ox = view->oy = -kAtlasOrigin / kAtlasCell;
    view->w = kAtlasDim, view->h = kAtlasDim;
}

}  // namespace

// 3 mm cubes over +-768 mm: 537 MB per cached plan (32-bit texels) + the brick pool (128 B per brick,
// one brick per 12 cubes: 1.4 GB).  Measured on the bench lattice
// with 8-bit texels: 4 mm / 384 is 2 % slower, 2.5 mm / 640 1 % faster.
int set_choice_volume_shape(float cell_mm, int dim) {
    if (!(cell_mm >= 0.5f && cell_mm <= 64.f) || dim < 16 || dim > 1024 || dim % 4 != 0) return -1;
    g_vol_cell.store(cell_mm), g_vol_dim.store(dim);
    return 0;
}
void get_choice_volume_shape(float* cell_mm, int* dim) { *cell_mm = g_vol_cell.load(), *dim = g_vol_dim.load(); }
int set_volume_bricks(int on) { return g_vol_bricks.exchange(on ? 1 : 0); }
void get_brick_stats(unsigned* used, unsigned* capacity) { *used = g_last_bricks.load(), *capacity = g_last_brick_cap.load(); }
unsigned long long table_builds() { return g_builds.load(); }
unsigned long long volume_builds_done() { return g_vol_builds_done.load(); }

bool tables_cached(const LegPlan& plan) {
    std::lock_guard<std::mutex> lock(g_mutex);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
    for (const Entry* c : g_cache[dev])
        if (c->used && std::memcmp(&c->plan, &plan, sizeof(LegPlan)) == 0) return true;
    return false;
}

cudaError_t acquire_tables(const LegPlan& plan, cudaStream_t stream, AtlasView* view, FastTables* tables,
                           TableLease* lease) {
    lease->entry = nullptr;
    std::lock_guard<std::mutex> lock(g_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    std::vector<Entry*>& cache = g_cache[dev];
    Entry* hit = nullptr;
    for (Entry* c : cache)
        if (c->used && std::memcmp(&c->plan, &plan, sizeof(LegPlan)) == 0) hit = c;
    if (!hit) {
        // victim: an unused entry, else the least recently used unpinned one, else a new entry
        Entry* victim = nullptr;
        for (Entry* c : cache)
            if (!c->used && c->pins == 0) victim = c;
        if (!victim && (int)cache.size() < kCacheEntries) {
            victim = new Entry();
            victim->device = dev;
            cache.push_back(victim);
        }
        if (!victim)
            for (Entry* c : cache)
                if (c->pins == 0 && (!victim || c->stamp < victim->stamp)) victim = c;
        if (!victim) {  // everything pinned by concurrent callers: spill over
            victim = new Entry();
            victim->device = dev;
            cache.push_back(victim);
        }
        victim->used = false;
        settle_volume(victim);
        victim->vol_ready = false;  // the volume belongs to the evicted plan: rebuilt on request
        if (!victim->cells) {
            e = allocate(victim);
            if (e != cudaSuccess) {
                release(victim);
                return e;
            }
        }
        e = order_after_uses(victim, stream);
        if (e != cudaSuccess) return e;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        atlas_build_kernel<<<sms * 8, 256, 0, stream>>>(plan, victim->cells, victim->linear, kAtlasDim,
                                                        kAtlasOrigin, kAtlasCell);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = cudaMemcpy2DToArrayAsync(victim->array, 0, 0, victim->linear, kAtlasDim, kAtlasDim, kAtlasDim,
                                     cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
        // other streams wait for this event on the device; the host does not wait
        e = cudaEventRecord(victim->ready, stream);
        if (e != cudaSuccess) return e;
        victim->ready_stream = stream;
        build_fast_tables(plan, &victim->tables);
        victim->used = true;
        std::memcpy(&victim->plan, &plan, sizeof(LegPlan));
        g_builds.fetch_add(1);
        hit = victim;
    }
    if (hit->ready_stream != stream) {
        e = cudaStreamWaitEvent(stream, hit->ready, 0);
        if (e != cudaSuccess) return e;
    }
    hit->stamp = ++g_clock;
    hit->pins++;
    lease->entry = hit;
    if (tables) *tables = hit->tables;
    fill_view(hit, view);
    return cudaSuccess;
}

void release_tables(TableLease* lease, cudaStream_t stream) {
    Entry* c = static_cast<Entry*>(lease->entry);
    if (!c) return;
    lease->entry = nullptr;
    std::lock_guard<std::mutex> lock(g_mutex);
    // remember that `stream` used the entry up to here
    int slot = -1;
    for (int k = 0; k < kUseSlots; k++)
        if (c->use_valid[k] && c->use_stream[k] == stream) slot = k;
    if (slot < 0)
        for (int k = 0; k < kUseSlots; k++)
            if (!c->use_valid[k]) slot = k;
    if (slot < 0 || cudaEventRecord(c->use_done[slot], stream) != cudaSuccess) {
        c->use_overflow = true;
        (void)cudaGetLastError();
    } else {
        c->use_stream[slot] = stream, c->use_valid[slot] = true;
    }
    c->pins--;
}

// wait = false: a volume that is not built yet is built in the BACKGROUND, on the entry's own
// stream, and cudaErrorNotReady is returned — the caller runs the two-tier sweep meanwhile (same
// results, bit for bit), so a one-off call never waits for a table it would use once; calls that
// come after the build has finished get the volume.  wait = true blocks until it is there.
cudaError_t get_choice_volume(const TableLease& lease, cudaStream_t stream, VolumeView* view, bool wait) {
    Entry* hit = static_cast<Entry*>(lease.entry);
    if (!hit) return cudaErrorInvalidValue;  // acquire_tables first
    std::lock_guard<std::mutex> lock(g_mutex);
    cudaError_t e;
    const int dim = g_vol_dim.load();
    const float cell = g_vol_cell.load();
    const int bricks_on = g_vol_bricks.load();
    if (hit->vol_ready && (hit->vol_dim != dim || hit->vol_cell != cell || hit->vol_bricks_on != bricks_on))
        hit->vol_ready = false;  // shape changed
    if (!hit->vol_ready && !hit->vol_building) {
        if (hit->vol_array && hit->vol_dim != dim) {
            // the old array may still be read by sweeps in flight
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return e;
            release_volume(hit);
        }
        if (!hit->vol_stream) {
            e = cudaStreamCreateWithFlags(&hit->vol_stream, cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
            e = cudaEventCreateWithFlags(&hit->vol_done, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        if (!hit->vol_array) {
            const cudaChannelFormatDesc fmt = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindUnsigned);
            e = cudaMalloc3DArray(&hit->vol_array, &fmt, make_cudaExtent(dim, dim, dim));
            if (e != cudaSuccess) return e;
            cudaResourceDesc res;
            std::memset(&res, 0, sizeof res);
            res.resType = cudaResourceTypeArray;
            res.res.array.array = hit->vol_array;
            cudaTextureDesc td;
            std::memset(&td, 0, sizeof td);
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;  // off the volume -> 0
            td.filterMode = cudaFilterModePoint;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            e = cudaCreateTextureObject(&hit->vol_tex, &res, &td, nullptr);
            if (e != cudaSuccess) {
                release_volume(hit);
                return e;
            }
            hit->vol_dim = dim;
        }
        if (bricks_on && !hit->vol_bricks) {
            // brick pool: one brick per 12 cubes (the shells around the decision surfaces of a leg
            // hold about 8 % of the cubes); without the memory the volume works without bricks
            hit->vol_brick_cap = (unsigned)((size_t)dim * dim * dim / 12);
            if (cudaMalloc((void**)&hit->vol_bricks, (size_t)hit->vol_brick_cap * 64 * sizeof(unsigned short)) != cudaSuccess ||
                cudaMalloc((void**)&hit->vol_brick_count, sizeof(unsigned)) != cudaSuccess) {
                (void)cudaGetLastError();
                if (hit->vol_bricks) cudaFree(hit->vol_bricks);
                hit->vol_bricks = nullptr, hit->vol_brick_cap = 0;
            }
        }
        // a sweep of the evicted plan may still read the array that is about to be rewritten, and
        // the build reads the atlas: order the build stream behind both
        e = order_after_uses(hit, hit->vol_stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(hit->vol_stream, hit->ready, 0);
        if (e != cudaSuccess) return e;
        const size_t cubes = (size_t)dim * dim * dim;
        // stream-ordered allocation on the build's own stream: freeing it later (cudaFreeAsync) does
        // not synchronise the device the way cudaFree does
        e = cudaMallocAsync((void**)&hit->vol_linear, cubes * sizeof(unsigned) + cubes / 64 + 64, hit->vol_stream);
        if (e != cudaSuccess) return e;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, hit->device);
        // the block map lives behind the texels in the same staging allocation
        unsigned char* block_done = reinterpret_cast<unsigned char*>(hit->vol_linear + cubes);
        volume_coarse_kernel<<<sms * kVolBuildCtasPerSm, 128, 0, hit->vol_stream>>>(hit->plan, hit->tables, hit->vol_linear, block_done,
                                                                  dim, cell);
        AtlasView atlas{};
        fill_view(hit, &atlas);
        volume_build_kernel<<<sms * kVolBuildCtasPerSm, 128, 0, hit->vol_stream>>>(hit->plan, hit->tables, atlas, hit->vol_linear,
                                                                   block_done, dim, cell);
        hit->vol_bricks_on = bricks_on;
        if (hit->vol_brick_count) cudaMemsetAsync(hit->vol_brick_count, 0, sizeof(unsigned), hit->vol_stream);
        if (bricks_on && hit->vol_bricks)
            brick_build_kernel<<<sms * kVolBuildCtasPerSm, 128, 0, hit->vol_stream>>>(
                hit->plan, hit->tables, atlas, hit->vol_linear, hit->vol_bricks, hit->vol_brick_count, hit->vol_brick_cap,
                dim, cell);
        e = cudaGetLastError();
        if (e == cudaSuccess) {
            cudaMemcpy3DParms cp;
            std::memset(&cp, 0, sizeof cp);
            cp.srcPtr = make_cudaPitchedPtr(hit->vol_linear, (size_t)dim * sizeof(unsigned), (size_t)dim, (size_t)dim);
            cp.dstArray = hit->vol_array;
            cp.extent = make_cudaExtent(dim, dim, dim);
            cp.kind = cudaMemcpyDeviceToDevice;
            e = cudaMemcpy3DAsync(&cp, hit->vol_stream);
        }
        if (e == cudaSuccess) e = cudaEventRecord(hit->vol_done, hit->vol_stream);
        if (e != cudaSuccess) {
            cudaFreeAsync(hit->vol_linear, hit->vol_stream);
            cudaStreamSynchronize(hit->vol_stream);
            hit->vol_linear = nullptr;
            return e;
        }
        hit->vol_cell = cell;
        hit->vol_building = true;
    }
    if (hit->vol_building) {
        e = wait ? cudaEventSynchronize(hit->vol_done) : cudaEventQuery(hit->vol_done);
        if (e == cudaErrorNotReady) {
            (void)cudaGetLastError();
            return cudaErrorNotReady;
        }
        if (e != cudaSuccess) return e;
        hit->vol_building = false;
        hit->vol_ready = true;
        g_vol_builds_done.fetch_add(1);
        cudaFreeAsync(hit->vol_linear, hit->vol_stream);
        hit->vol_linear = nullptr;
        if (hit->vol_brick_count) {  // diagnostics (lrm_get_stat): the build has finished, the copy does not wait
            unsigned used = 0;
            if (cudaMemcpy(&used, hit->vol_brick_count, sizeof used, cudaMemcpyDeviceToHost) == cudaSuccess)
                g_last_bricks.store(used), g_last_brick_cap.store(hit->vol_brick_cap);
            else
                (void)cudaGetLastError();
        }
    }
    (void)stream;  // vol_done has completed on the host's clock: no device-side wait needed
    view->tex = hit->vol_tex;
    view->inv_cell = 1.0f / hit->vol_cell;
    view->o = 0.5f * (float)hit->vol_dim;
    view->oy = view->o + kVolShiftY;
    view->dim = hit->vol_dim;
    view->bricks = (hit->vol_bricks_on && hit->vol_bricks) ? hit->vol_bricks : nullptr;
    return cudaSuccess;
}

}  // namespace lrm
