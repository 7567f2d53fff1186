// bulk_copy.cuh — thin inline-PTX wrappers for the sm_100a asynchronous bulk-copy engine (TMA in
// its 1-D "bulk" form, SASS UBLKCP) and the mbarrier that tracks its completion.  The streaming
// kernels stage contiguous slabs of points global -> shared -> global with these instead of
// per-thread LDG/STG: one elected thread issues a multi-KB copy, the data path never touches the
// register file, and many KB per SM stay in flight without costing occupancy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lrm {
namespace bulk {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals)
                 : "memory");
}
// make the initialised barrier visible to the async proxy before the first copy targets it
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order generic-proxy shared-memory writes before subsequent async-proxy (bulk copy) reads
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on `bar`.  dst/src 16-B aligned, bytes % 16 == 0.
__device__ __forceinline__ void load(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                     uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
// shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_addr(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void commit_group() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until at most N of this thread's store groups are still READING shared memory
template <int N>
__device__ __forceinline__ void wait_group_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// wait until at most N store groups are incomplete (writes to global performed)
template <int N>
__device__ __forceinline__ void wait_group() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace bulk
}  // namespace lrm
