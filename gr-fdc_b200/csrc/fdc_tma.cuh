/* fdc_tma.cuh -- bulk asynchronous copies global -> shared memory (TMA engine, cp.async.bulk) completing on an
 * mbarrier: the "TMA-staged tile" used by the 32-point extract, which has no registers left for a prefetch tile.
 * One elected thread arms the barrier with the byte count and issues one copy per channel slice; the data lands in
 * shared memory while the CTA computes the previous tile; every thread then waits on the barrier's phase bit.
 * Device only (sm_90+); under the host emulator the same staging is a memcpy (tests/emu). */
#ifndef FDC_TMA_CUH
#define FDC_TMA_CUH
#if defined(__CUDACC__)
#include <stdint.h>

namespace fdc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/* one arrival + the number of bytes the copies issued next will deliver */
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* src and dst 16-byte aligned, bytes a multiple of 16 */
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* ask the TMA engine to bring [src, src + bytes) into L2 (no destination, no completion to wait for): one instruction of one
 * thread per contiguous run; src 16-byte aligned, bytes a multiple of 16 */
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_global, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_global), "r"(bytes) : "memory");
}
/* wait until the barrier's phase with the given parity has completed; traps instead of hanging the GPU if the bytes never
 * arrive (a mis-sized expect_tx would otherwise spin forever) */
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (clock64() - t0 > (1LL << 32)) __trap();          /* ~2 s at 2 GHz */
    }
}

}  // namespace fdc
#endif
#endif
