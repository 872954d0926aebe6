/* fdc_kcommon.cuh -- shared bits of the kernel translation units (device only). */
#ifndef FDC_KCOMMON_CUH
#define FDC_KCOMMON_CUH
#include "fdc_launch.h"
#include "fdc_tile_fft.cuh"

namespace fdc {

extern __shared__ __align__(16) unsigned char fdc_smem_raw[];

/* resident CTAs per SM the register allocator should aim for.  Prefetching kernels hold two register tiles
 * (32 + 32 floats) and get 128 registers: 2 CTAs of 256 threads or 1 of 512; kernels without prefetch 3 / 2 / 1. */
constexpr int min_ctas(int threads, bool prefetch)
{
    return prefetch ? (threads <= 128 ? 4 : (threads <= 256 ? 2 : 1)) : (threads <= 256 ? 3 : (threads <= 512 ? 2 : 1));
}
/* a 1024-thread CTA has 64 registers per thread: no room for a second register tile */
constexpr bool can_prefetch(int threads) { return threads <= 512; }

/* tile batch for a transform length: 4096-point tiles (256 threads), one signal per CTA above that */
constexpr int tile_batch(int L) { return L >= 4096 ? 1 : 4096 / L; }

/* dynamic shared memory of a tile kernel: exchange buffer + (small lengths) the twiddle table */
template <class ENG> constexpr size_t tile_smem_bytes() { return sizeof(float2) * (size_t)(ENG::SMEM_ELEMS + tw_smem_elems(ENG::L)); }

/* column / row tiles of the four-step kernels */
constexpr int big_tile_batch(int L) { return L <= 256 ? 16 : 4096 / L; }

#define FDC_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

/* Persistent launch geometry: as many CTAs as are resident at once (occupancy x SM count), never more than there are
 * tiles; `multiple` > 1 rounds down to a multiple (a CTA then keeps the same inner tile index for its whole life). */
/* resident CTAs of `kernel` on the current device (cached per device and kernel entry point; also opts the kernel in to
 * its dynamic shared memory size).  Defined in fdc_k_misc.cu. */
cudaError_t kernel_capacity(const void* kernel, int threads, size_t smem, int* capacity);
template <class K> cudaError_t persistent_grid(K kernel, int threads, size_t smem, long ntiles, int multiple, unsigned* grid)
{
    int capacity = 0;
    FDC_CHECK(kernel_capacity((const void*)kernel, threads, smem, &capacity));
    long g = capacity;
    if (multiple > 1 && g >= multiple) g -= g % multiple;
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    *grid = (unsigned)g;
    return cudaSuccess;
}

/* the body shared by all tile kernels: persistent loop with or without register prefetch */
template <class ENG, bool PF, class Tiles>
__device__ __forceinline__ void tile_kernel_body(const Tiles& tiles, const float2* tw, long ntiles)
{
    tile_fft_loop<ENG, PF>(reinterpret_cast<float2*>(fdc_smem_raw), tw, tiles, (long)blockIdx.x, (long)gridDim.x, ntiles);
}

}  // namespace fdc
#endif
