// cell_grid.cuh — xy bucket grid over a point cloud (cell-sorted float4 points + per-cell z range),
// built on the device with a counting sort.  Shared by the positionability search and the
// body-space octree.  Everything is internal-linkage: each translation unit gets its own copy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <vector>

namespace lrm {
namespace {

struct CellGrid {
    float x0, y0, inv_cell;
    int nx, ny;
    const int* cell_start;   // nx*ny + 1
    const float4* pts;       // cell-sorted (x, y, z, keep)
    const float2* cell_z;    // per cell (zmin, zmax)
    const float2* cell_ball; // per cell (z of the centre, radius of the bounding ball); radius < 0: empty cell
    const float2* blk_ball;  // the same per block of 4 x 4 cells (nbx x nby blocks, aligned to cell 0)
    int nbx, nby;
    int n;
};

// ---- grid construction -------------------------------------------------------------------------
__global__ void bounds_kernel(const float* __restrict__ xyz, size_t n, float* out4) {
    // out4 = {xmin, ymin, -xmax, -ymax} as atomicMin on ordered ints
    float xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = xyz[3 * i], y = xyz[3 * i + 1];
        xmin = fminf(xmin, x), xmax = fmaxf(xmax, x), ymin = fminf(ymin, y), ymax = fmaxf(ymax, y);
    }
    for (int o = 16; o; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        auto enc = [](float f) {  // order-preserving float -> int
            int i = __float_as_int(f);
            return i >= 0 ? i : i ^ 0x7fffffff;
        };
        atomicMin(reinterpret_cast<int*>(out4) + 0, enc(xmin));
        atomicMin(reinterpret_cast<int*>(out4) + 1, enc(ymin));
        atomicMin(reinterpret_cast<int*>(out4) + 2, enc(-xmax));
        atomicMin(reinterpret_cast<int*>(out4) + 3, enc(-ymax));
    }
}

__device__ __forceinline__ int cell_of(const CellGrid& g, float x, float y) {
    int cx = (int)((x - g.x0) * g.inv_cell), cy = (int)((y - g.y0) * g.inv_cell);
    cx = min(max(cx, 0), g.nx - 1), cy = min(max(cy, 0), g.ny - 1);
    return cy * g.nx + cx;
}

__global__ void count_kernel(CellGrid g, const float* __restrict__ xyz, int* __restrict__ counts,
                             int* __restrict__ zmin, int* __restrict__ zmax) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)g.n; i += stride) {
        const int c = cell_of(g, xyz[3 * i], xyz[3 * i + 1]);
        atomicAdd(&counts[c], 1);
        int zi = __float_as_int(xyz[3 * i + 2]);
        zi = zi >= 0 ? zi : zi ^ 0x7fffffff;
        atomicMin(&zmin[c], zi);
        atomicMax(&zmax[c], zi);
    }
}

// single-CTA exclusive scan (cell counts are a few 1e4..1e6 entries; setup cost only)
__global__ void scan_kernel(const int* __restrict__ counts, int* __restrict__ start, int ncell) {
    __shared__ int carry;
    __shared__ int warp_sum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int v = i < ncell ? counts[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = threadIdx.x < (blockDim.x >> 5) ? warp_sum[threadIdx.x] : 0;
            int wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            warp_sum[threadIdx.x] = wi - w;  // exclusive prefix of warp sums
        }
        __syncthreads();
        const int excl = carry + warp_sum[threadIdx.x >> 5] + incl - v;
        if (i < ncell) start[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[ncell] = carry;
}

__global__ void scatter_kernel(CellGrid g, const float* __restrict__ xyz,
                               const uint8_t* __restrict__ keep, int* __restrict__ cursor,
                               float4* __restrict__ sorted, const int* __restrict__ zmin,
                               const int* __restrict__ zmax, float2* __restrict__ cell_z,
                               float2* __restrict__ cell_ball, const int* __restrict__ counts) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = tid; i < (size_t)g.n; i += stride) {
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        const int c = cell_of(g, x, y);
        const int slot = atomicAdd(&cursor[c], 1);
        sorted[slot] = make_float4(x, y, z, (keep == nullptr || keep[i]) ? 1.f : 0.f);
    }
    const int ncell = g.nx * g.ny;
    for (size_t c = tid; c < (size_t)ncell; c += stride) {
        auto dec = [](int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); };
        const float z0 = dec(zmin[c]), z1 = dec(zmax[c]);
        cell_z[c] = make_float2(z0, z1);
        // bounding ball of the cell's points: half diagonal of the cell (+0.2 % for points binned
        // across an edge by rounding) and half the z range
        const float cell = 1.0f / g.inv_cell, hz = 0.5f * (z1 - z0);
        cell_ball[c] = counts[c] > 0 ? make_float2(0.5f * (z0 + z1), sqrtf(0.5f * cell * cell * 1.004f + hz * hz) + 1.0e-3f)
                                     : make_float2(0.f, -1.f);
    }
}

// bounding ball of every block of 4 x 4 cells (two-level cell pruning of the positionability search)
__global__ void block_ball_kernel(CellGrid g, const int* __restrict__ counts, const float2* __restrict__ cell_z,
                                  float2* __restrict__ blk_ball) {
    const int nblk = g.nbx * g.nby;
    const float cell = 1.0f / g.inv_cell;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += gridDim.x * blockDim.x) {
        const int bx = b % g.nbx, by = b / g.nbx;
        float z0 = INFINITY, z1 = -INFINITY;
        for (int j = 0; j < 16; j++) {
            const int cx = 4 * bx + (j & 3), cy = 4 * by + (j >> 2);
            if (cx < g.nx && cy < g.ny && counts[cy * g.nx + cx] > 0) {
                const float2 zr = cell_z[cy * g.nx + cx];
                z0 = fminf(z0, zr.x), z1 = fmaxf(z1, zr.y);
            }
        }
        if (z1 >= z0) {
            // half diagonal of the 4-cell square (+0.2 % as for the cells) and half the z range
            const float hz = 0.5f * (z1 - z0);
            blk_ball[b] = make_float2(0.5f * (z0 + z1), sqrtf(8.0f * cell * cell * 1.004f + hz * hz) + 1.0e-3f);
        } else {
            blk_ball[b] = make_float2(0.f, -1.f);
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct DevBuf {
    std::vector<void*> ptrs;
    template <class T>
    cudaError_t alloc(T** out, size_t count) {
        cudaError_t e = cudaMalloc((void**)out, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
    ~DevBuf() {
        for (void* p : ptrs) cudaFree(p);
    }
};

#define POSIT_CHECK(call)                  \
    do {                                   \
        cudaError_t e_ = (call);           \
        if (e_ != cudaSuccess) return e_;  \
    } while (0)

cudaError_t build_grid(DevBuf& mem, const float* xyz, size_t n, const uint8_t* keep, float cell,
                       cudaStream_t stream, CellGrid* out) {
    const int enc_inf = 0x7f800000;
    float* d_bounds;
    POSIT_CHECK(mem.alloc(&d_bounds, 4));
    int init[4] = {enc_inf, enc_inf, enc_inf, enc_inf};
    POSIT_CHECK(cudaMemcpyAsync(d_bounds, init, sizeof init, cudaMemcpyHostToDevice, stream));
    if (n) bounds_kernel<<<296, 256, 0, stream>>>(xyz, n, d_bounds);
    int h[4];
    POSIT_CHECK(cudaMemcpyAsync(h, d_bounds, sizeof h, cudaMemcpyDeviceToHost, stream));
    POSIT_CHECK(cudaStreamSynchronize(stream));
    auto dec = [](int i) {
        i = i >= 0 ? i : i ^ 0x7fffffff;
        float f;
        memcpy(&f, &i, 4);
        return f;
    };
    float xmin = dec(h[0]), ymin = dec(h[1]), xmax = -dec(h[2]), ymax = -dec(h[3]);
    if (n == 0 || !(xmax >= xmin) || !(ymax >= ymin)) xmin = ymin = 0.f, xmax = ymax = 1.f;
    if (cell <= 0.f) {
        // automatic: about 48 points per cell, between 32 and 256 mm
        const double area = std::fmax(1.0, (double)(xmax - xmin)) * std::fmax(1.0, (double)(ymax - ymin));
        cell = (float)std::sqrt(area * 48.0 / (double)(n ? n : 1));
        cell = std::fmin(256.f, std::fmax(32.f, cell));
    }
    // keep the cell table bounded (L2-friendly) whatever the map extent
    while (((double)(xmax - xmin) / cell + 1) * ((double)(ymax - ymin) / cell + 1) > 4.0e6) cell *= 2;
    CellGrid g;
    g.x0 = xmin, g.y0 = ymin, g.inv_cell = 1.0f / cell;
    g.nx = (int)((xmax - xmin) * g.inv_cell) + 1;
    g.ny = (int)((ymax - ymin) * g.inv_cell) + 1;
    g.n = (int)n;
    const int ncell = g.nx * g.ny;
    int *counts, *start, *zmin, *zmax, *cursor;
    float4* sorted;
    float2 *cell_z, *cell_ball, *blk_ball;
    g.nbx = (g.nx + 3) / 4, g.nby = (g.ny + 3) / 4;
    POSIT_CHECK(mem.alloc(&blk_ball, (size_t)g.nbx * g.nby));
    POSIT_CHECK(mem.alloc(&counts, ncell));
    POSIT_CHECK(mem.alloc(&start, ncell + 1));
    POSIT_CHECK(mem.alloc(&cursor, ncell));
    POSIT_CHECK(mem.alloc(&zmin, ncell));
    POSIT_CHECK(mem.alloc(&zmax, ncell));
    POSIT_CHECK(mem.alloc(&sorted, n));
    POSIT_CHECK(mem.alloc(&cell_z, ncell));
    POSIT_CHECK(mem.alloc(&cell_ball, ncell));
    POSIT_CHECK(cudaMemsetAsync(counts, 0, ncell * sizeof(int), stream));
    POSIT_CHECK(cudaMemsetAsync(zmin, 0x7f, ncell * sizeof(int), stream));  // large positive
    POSIT_CHECK(cudaMemsetAsync(zmax, 0x80, ncell * sizeof(int), stream));  // large negative
    if (n) count_kernel<<<592, 256, 0, stream>>>(g, xyz, counts, zmin, zmax);
    scan_kernel<<<1, 1024, 0, stream>>>(counts, start, ncell);
    POSIT_CHECK(cudaMemcpyAsync(cursor, start, ncell * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    scatter_kernel<<<592, 256, 0, stream>>>(g, xyz, keep, cursor, sorted, zmin, zmax, cell_z, cell_ball, counts);
    block_ball_kernel<<<296, 256, 0, stream>>>(g, counts, cell_z, blk_ball);
    POSIT_CHECK(cudaGetLastError());
    g.cell_start = start, g.pts = sorted, g.cell_z = cell_z, g.cell_ball = cell_ball, g.blk_ball = blk_ball;
    *out = g;
    return cudaSuccess;
}

}  // namespace
}  // namespace lrm
