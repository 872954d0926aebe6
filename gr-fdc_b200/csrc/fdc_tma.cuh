/* fdc_tma.cuh -- the bulk (TMA engine) instructions the kernels use: cp.async.bulk.prefetch.L2 pulls the next tile's
 * operands into L2 with ONE instruction of one thread (no registers, no completion to wait for) while the CTA transforms the
 * current tile; the mbarrier / cp.async.bulk global -> shared helpers are kept for staging experiments (a TMA-staged
 * extract tile was measured at 65 against 80 Gsample/s for the register-staged one on cfg4 and removed,
 * profiles/r2_sweep_engine_variants_packed.txt).  Device only (sm_90+). */
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
