// plane_atlas.cu — builds and caches the plane atlas of a LegPlan (see leg_math.cuh).
//
// One byte per cell of a regular grid over the femur plane: the certified outcome of plane_clamp
// (valid bit, sector, winning candidate) or "impure".  Built on the device by probing every cell
// centre with the instrumented evaluation (plane_probe): a cell is certified only when the smallest
// decision margin exceeds the cell's half diagonal plus a float-rounding allowance (kAtlasNeed*).  The atlas is a
// pure accelerator: the streaming kernel falls back to the full evaluation for every point it
// cannot certify.  16 MiB per (leg, orientation); a small per-device LRU keeps the last few.
#include <cuda_runtime.h>

#include <cstring>
#include <mutex>

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

struct Entry {
    bool used = false;
    int device = -1;
    unsigned long long stamp = 0;
    LegPlan plan;
    unsigned char* cells = nullptr;   // 8 x 4 blocked
    unsigned char* linear = nullptr;  // row-major staging copy for the texture upload
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
    FastTables tables;  // yaw-sector table of the same plan (host-built, ~0.5 ms: cached with the atlas)
};
void release(Entry* c) {
    if (c->tex) cudaDestroyTextureObject(c->tex);
    if (c->array) cudaFreeArray(c->array);
    if (c->cells) cudaFree(c->cells);
    if (c->linear) cudaFree(c->linear);
    c->tex = 0, c->array = nullptr, c->cells = nullptr, c->linear = nullptr;
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
    return cudaCreateTextureObject(&c->tex, &res, &td, nullptr);
}
constexpr int kCacheEntries = 4;
Entry g_cache[kCacheEntries];
unsigned long long g_clock = 0;
std::mutex g_mutex;

}  // namespace

cudaError_t get_plane_atlas(const LegPlan& plan, cudaStream_t stream, AtlasView* view, FastTables* tables) {
    std::lock_guard<std::mutex> lock(g_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    Entry* hit = nullptr;
    Entry* victim = &g_cache[0];
    for (Entry& c : g_cache) {
        if (c.used && c.device == dev && std::memcmp(&c.plan, &plan, sizeof(LegPlan)) == 0) hit = &c;
        if (!c.used || c.stamp < victim->stamp || (victim->used && !c.used)) victim = &c;
    }
    if (!hit) {
        for (Entry& c : g_cache)
            if (!c.used) {
                victim = &c;
                break;
            }
        if (victim->used && victim->device != dev) {
            // evicting another device's atlas: free it on that device
            int cur = dev;
            cudaSetDevice(victim->device);
            release(victim);
            cudaSetDevice(cur);
        } else if (victim->used) {
            // reuse the buffers: earlier kernels on other streams may still read them
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return e;
        }
        victim->used = false;
        if (!victim->cells) {
            e = allocate(victim);
            if (e != cudaSuccess) {
                release(victim);
                return e;
            }
        }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        atlas_build_kernel<<<sms * 8, 256, 0, stream>>>(plan, victim->cells, victim->linear, kAtlasDim,
                                                        kAtlasOrigin, kAtlasCell);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = cudaMemcpy2DToArrayAsync(victim->array, 0, 0, victim->linear, kAtlasDim, kAtlasDim, kAtlasDim,
                                     cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
        // other streams may use this entry next: make the build visible device-wide
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        build_fast_tables(plan, &victim->tables);
        victim->used = true;
        victim->device = dev;
        std::memcpy(&victim->plan, &plan, sizeof(LegPlan));
        hit = victim;
    }
    hit->stamp = ++g_clock;
    if (tables) *tables = hit->tables;
    view->cells = hit->cells;
    view->tex = hit->tex;
    view->inv_cell = 1.0f / kAtlasCell;
    view->ox = view->oy = -kAtlasOrigin / kAtlasCell;
    view->w = kAtlasDim, view->h = kAtlasDim;
    return cudaSuccess;
}

}  // namespace lrm
