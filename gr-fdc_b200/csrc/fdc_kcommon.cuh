/* fdc_kcommon.cuh -- shared bits of the kernel translation units (device only). */
#ifndef FDC_KCOMMON_CUH
#define FDC_KCOMMON_CUH
#include "fdc_launch.h"
#include "fdc_tile_fft.cuh"

namespace fdc {

extern __shared__ __align__(16) unsigned char fdc_smem_raw[];

/* resident CTAs per SM the register allocator should aim for */
constexpr int min_ctas(int threads) { return threads <= 256 ? 3 : (threads <= 512 ? 2 : 1); }

/* tile batch for a transform length: 4096-point tiles (256 threads), one signal per CTA above that */
constexpr int tile_batch(int L) { return L >= 4096 ? 1 : 4096 / L; }

template <class K> cudaError_t set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}
#define FDC_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

}  // namespace fdc
#endif
