// uaisa_tc.cuh -- inline-PTX helpers for the 5th-generation tensor cores (tcgen05 / TMEM), mbarriers and the bulk-copy engine on sm_100a.
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uaisa {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: start>>4 | LBO(=1)<<16 | SBO(=1024>>4)<<32 | version 1 <<46 | layout SWIZZLE_128B(2) <<61.
// The tile is [rows][32 fp32] with 128-byte rows, 8-row groups 1024 bytes apart, 16-byte chunk k4 of row r stored at chunk (k4 ^ (r & 7)).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor for kind::tf32: D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), both K-major, N>>3 <<17, M>>4 <<24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }   // generic-proxy smem writes -> async proxy (MMA, bulk copies)

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(mbar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// bulk copy global -> shared by the copy engine (SASS UBLKCP); completes `bytes` transaction bytes on the mbarrier.  16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}

// The same copy issued by ONE CTA of a cluster, delivered to the CTA-relative address `sdst` (and signalled on the CTA-relative mbarrier `mbar`) of EVERY
// CTA in `cta_mask`: a weight tile read once from L2 feeds several SMs.
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t mbar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(sdst), "l"(gsrc), "r"(bytes),
                 "r"(mbar), "h"(cta_mask)
                 : "memory");
}
// tcgen05.commit that arrives on the mbarrier at the same CTA-relative offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t mbar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(mbar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory"); }

// 32 lanes x 16 consecutive columns of 32 bits: thread t of the warp gets row (lane base + t), v[i] = column (col base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
        "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
        "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    tmem_ld16(taddr, v);
    tmem_ld16(taddr + 16, v + 16);
}

// 16-byte chunk (row, k4) of a [rows x 32 fp32] K-major SWIZZLE_128B tile
__device__ __forceinline__ float4 *swz(char *tile_base, int row, int k4) {
    return reinterpret_cast<float4 *>(tile_base + (row >> 3) * 1024 + (row & 7) * 128 + ((k4 ^ (row & 7)) << 4));
}
__host__ __device__ __forceinline__ int swz_float_index(int row, int k) {      // float index of element (row, k) inside the tile image
    return (row >> 3) * 256 + (row & 7) * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3));
}

}  // namespace tc
}  // namespace uaisa
