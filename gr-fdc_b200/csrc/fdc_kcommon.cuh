/* fdc_kcommon.cuh -- shared bits of the kernel translation units (device only). */
#ifndef FDC_KCOMMON_CUH
#define FDC_KCOMMON_CUH
#include "fdc_launch.h"
#include "fdc_tile_fft.cuh"

namespace fdc {

extern __shared__ __align__(16) unsigned char fdc_smem_raw[];

/* resident CTAs per SM the register allocator should aim for.  Prefetching kernels hold two register tiles
 * (32 + 32 floats) and get 128 registers: 2 CTAs of 256 threads or 1 of 512; kernels without prefetch 4 / 2 / 1 (the 16-point register tile fits 64 registers). */
constexpr int min_ctas(int threads, bool prefetch, int L = 0)
{
    return prefetch ? (threads <= 128 ? 4 : (threads <= 256 ? 2 : 1))
                    : (L <= 1024 ? 4 : (threads <= 256 ? 2 : 1));       /* 3 passes fit 64 registers; 4 passes get 128 */
}
/* a 1024-thread CTA has 64 registers per thread: no room for a second register tile */
constexpr bool can_prefetch(int threads) { return threads <= 512; }

/* tile batch for a transform length: 4096-point tiles (256 threads), one signal per CTA above that */
constexpr int tile_batch(int L) { return L >= 4096 ? 1 : 4096 / L; }

/* dynamic shared memory of a tile kernel: exchange buffer + (small lengths) the twiddle table */
template <class ENG> constexpr size_t tile_smem_bytes() { return sizeof(float2) * (size_t)(ENG::SMEM_ELEMS + tw_smem_elems(ENG::L, ENG::E)); }

/* column / row tiles of the four-step kernels */
constexpr int big_tile_batch(int L) { return L < 256 ? 4096 / L : (L == 256 ? 16 : 4096 / L); }

#define FDC_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

/* Persistent launch geometry: as many CTAs as are resident at once (occupancy x SM count), never more than there are
 * tiles; `multiple` > 1 rounds down to a multiple (a CTA then keeps the same inner tile index for its whole life). */
/* resident CTAs of `kernel` on the current device (cached per device and kernel entry point; also opts the kernel in to
 * its dynamic shared memory size).  Defined in fdc_k_misc.cu. */
cudaError_t kernel_capacity(const void* kernel, int threads, size_t smem, int* capacity, int limit);
template <class K> cudaError_t persistent_grid(K kernel, int threads, size_t smem, long ntiles, int multiple, unsigned* grid, int limit = 0)
{
    int capacity = 0;
    FDC_CHECK(kernel_capacity((const void*)kernel, threads, smem, &capacity, limit));
    long g = capacity;
    /* `multiple`: the kernel relies on tile (blockIdx + k * grid) % multiple == blockIdx % multiple (a column CTA keeps its
     * four-step twiddle slice).  With fewer resident CTAs than that the grid is oversubscribed: a second wave, same result. */
    if (multiple > 1) g = g >= multiple ? g - g % multiple : multiple;
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    *grid = (unsigned)g;
    return cudaSuccess;
}

/* Launch of a tile kernel.  With programmatic dependent launch (sm_90+) the kernel may start while its predecessor in the
 * stream is still draining: its prologue (twiddle tables into shared memory) overlaps the predecessor's tail and the
 * launch latency; tile_fft_loop waits for the predecessor's results with cudaGridDependencySynchronize(). */
template <class... KA, class... A>
cudaError_t launch_tile_kernel(void (*kernel)(KA...), unsigned grid, int threads, size_t smem, cudaStream_t s, A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = tuning().pdl ? 1 : 0;
    count_launch();
    return cudaLaunchKernelEx(&cfg, kernel, KA(args)...);
}

/* the body shared by all tile kernels: persistent loop with or without register prefetch */
template <class ENG, bool PF, class Tiles, bool TW_SMEM = true>
__device__ __forceinline__ void tile_kernel_body(const Tiles& tiles, const float2* tw, long ntiles)
{
    tile_fft_loop<ENG, PF, Tiles, TW_SMEM>(reinterpret_cast<float2*>(fdc_smem_raw), tw, tiles, (long)blockIdx.x, (long)gridDim.x, ntiles);
}

}  // namespace fdc
#endif
