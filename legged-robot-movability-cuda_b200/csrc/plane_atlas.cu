// plane_atlas.cu — builds and caches the plane atlas of a LegPlan (see leg_math.cuh).
//
// One byte per cell of a regular grid over the femur plane: the certified outcome of plane_clamp
// (valid bit, sector, winning candidate) or "impure".  Built on the device by probing every cell
// centre with the instrumented evaluation (plane_probe): a cell is certified only when the smallest
// decision margin exceeds the cell's half diagonal plus a float-rounding allowance.  The atlas is a
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

__global__ void atlas_build_kernel(const __grid_constant__ LegPlan L, signed char* __restrict__ cells,
                                   int dim, float origin, float cell) {
    __shared__ SectorTable table;
    fill_sector_table(L, &table, threadIdx.x, blockDim.x);
    __syncthreads();
    const size_t total = (size_t)dim * dim;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // half diagonal + allowance for float rounding of the margins and of the cell index
    const float need = cell * 0.70711f * 1.02f + 2.0e-3f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ix = (int)(i % dim), iy = (int)(i / dim);
        const float X = origin + ((float)ix + 0.5f) * cell, Y = origin + ((float)iy + 0.5f) * cell;
        const PlaneProbe pr = plane_probe(L, table, X, Y);
        cells[atlas_index(dim, ix, iy)] = (signed char)(pr.safety > need ? pr.label : 0x80);
    }
}

struct Entry {
    bool used = false;
    int device = -1;
    unsigned long long stamp = 0;
    LegPlan plan;
    signed char* cells = nullptr;
};
constexpr int kCacheEntries = 4;
Entry g_cache[kCacheEntries];
unsigned long long g_clock = 0;
std::mutex g_mutex;

}  // namespace

cudaError_t get_plane_atlas(const LegPlan& plan, cudaStream_t stream, AtlasView* view) {
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
            cudaFree(victim->cells);
            cudaSetDevice(cur);
            victim->cells = nullptr;
        } else if (victim->used) {
            // reuse the buffer: earlier kernels on other streams may still read it
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return e;
        }
        if (!victim->cells) {
            e = cudaMalloc((void**)&victim->cells, (size_t)kAtlasDim * kAtlasDim);
            if (e != cudaSuccess) {
                victim->used = false;
                return e;
            }
        }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        atlas_build_kernel<<<sms * 8, 256, 0, stream>>>(plan, victim->cells, kAtlasDim, kAtlasOrigin,
                                                        kAtlasCell);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        // other streams may use this entry next: make the build visible device-wide
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        victim->used = true;
        victim->device = dev;
        std::memcpy(&victim->plan, &plan, sizeof(LegPlan));
        hit = victim;
    }
    hit->stamp = ++g_clock;
    view->cells = hit->cells;
    view->x0 = kAtlasOrigin, view->y0 = kAtlasOrigin;
    view->inv_cell = 1.0f / kAtlasCell;
    view->w = kAtlasDim, view->h = kAtlasDim;
    return cudaSuccess;
}

}  // namespace lrm
