/* fdc_k_extract.cu -- K2: batched channel extraction (bin gather, table multiply, half swap, backward FFT,
 * overlap discard, gain), all channel-blocks of one slice length in one launch.  Replaces
 * vector_cut_vxx + phase_shifting_windowing_vcc (lib/vector_cut_vxx_impl.cc:59-72,
 * lib/phase_shifting_windowing_vcc_impl.cc:72-86) with the third-party inverse fft_vcc / multiply_const stages
 * between and behind them, and process_channel of the activity-gated blocks. */
#include "fdc_kcommon.cuh"

#ifndef FDC_EXT32_CTAS
#define FDC_EXT32_CTAS 5
#endif

namespace fdc {

template <int L, int B, bool PF>
__global__ void __launch_bounds__((TileFFT<L, B, -1, false, false>::T), min_ctas(TileFFT<L, B, -1, false, false>::T, PF, L))
k_extract(const ExtractParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false>, PF>(ExtractTiles<L, B>{p}, tw, ntiles);
}
/* packed tiles (few channels: a tile spans several blocks, fdc_functors.cuh) */
template <int L, int B>
__global__ void __launch_bounds__((TileFFT<L, B, -1, false, false>::T), min_ctas(TileFFT<L, B, -1, false, false>::T, false, L))
k_extract_packed(const ExtractParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false>, false>(PackedExtractTiles<L, B>{p}, tw, ntiles);
}
template <int L, int B>
__global__ void __launch_bounds__(256, 4)
k_extract8_packed(const ExtractParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false, 8>, false>(PackedExtractTiles<L, B>{p}, tw, ntiles);
}
template <int L, int B>
__global__ void __launch_bounds__(128, 5)
k_extract32_packed(const ExtractParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false, 32>, false>(PackedExtractTiles<L, B>{p}, tw, ntiles);
}
/* 8 points per thread: tiles of 2048 points on 256 threads at <= 64 registers, four CTAs (32 warps) per SM instead of two
 * CTAs (16 warps) -- the extract is latency bound, not bandwidth bound, so it is the warps in flight that count */
template <int L, int B, bool PF>
__global__ void __launch_bounds__(256, 4)
k_extract8(const ExtractParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false, 8>, PF>(ExtractTiles<L, B>{p}, tw, ntiles);
}
/* 32 points per thread (512 = 16 * 32, 1024 = 32 * 32): one shared-memory exchange instead of two; the extract is limited by
 * the L1/shared-memory pipe, so the exchange it does not do is the saving.  128 threads at 96 registers, 5 CTAs per SM (measured: 4 CTAs / 127 registers
 * 79.1 Gsample/s, 5 CTAs 80.4, 6 CTAs with the twiddles left in global memory to fit the shared memory 69). */
template <int L, int B>
__global__ void __launch_bounds__(128, FDC_EXT32_CTAS)
k_extract32(const ExtractParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false, 32>, false>(ExtractTiles<L, B>{p}, tw, ntiles);
}
template <int L, int B>
__global__ void __launch_bounds__((TileFFT<L, B, -1, false, false>::T), min_ctas(TileFFT<L, B, -1, false, false>::T, false, L))
k_jobs(const JobParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false>, false>(JobTiles<L, B>{p}, tw, ntiles);
}

/* fewer than half a tile of channels and more than one block: pack several blocks into a tile */
static bool use_packed(ExtractParams& p, int B, long* ntiles)
{
    p.bpt = 1;
    if (B < 2 || p.nsel * 2 > B || p.nb < 2 || !tuning().pack) return false;
    p.bpt = B / p.nsel; p.ny = 1;
    *ntiles = (p.nb + p.bpt - 1) / p.bpt;
    return true;
}
template <int L, bool PF> static cudaError_t go_extract(const ExtractParams& p0, cudaStream_t s)
{
    constexpr int B = tile_batch(L);
    typedef TileFFT<L, B, -1, false, false> ENG;
    ExtractParams p = p0;
    if constexpr (B >= 2) {
        long nt = 0;
        if (use_packed(p, B, &nt)) {
            unsigned g = 1;
            FDC_CHECK(persistent_grid(k_extract_packed<L, B>, ENG::T, tile_smem_bytes<ENG>(), nt, 1, &g, tuning().ctas_ext));
            return launch_tile_kernel(k_extract_packed<L, B>, g, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L), nt);
        }
    }
    p.ny = (p.nsel + B - 1) / B;
    const long ntiles = p.nb * p.ny;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_extract<L, B, PF>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid, tuning().ctas_ext));
    return launch_tile_kernel(k_extract<L, B, PF>, grid, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L), ntiles);
}
template <int L, bool PF> static cudaError_t go_extract8(const ExtractParams& p0, cudaStream_t s)
{
    constexpr int B = 2048 / L;
    typedef TileFFT<L, B, -1, false, false, 8> ENG;
    static_assert(ENG::T == 256, "2048-point tiles");
    ExtractParams p = p0;
    {
        long nt = 0;
        if (use_packed(p, B, &nt)) {
            unsigned g = 1;
            FDC_CHECK(persistent_grid(k_extract8_packed<L, B>, ENG::T, tile_smem_bytes<ENG>(), nt, 1, &g));
            return launch_tile_kernel(k_extract8_packed<L, B>, g, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L, 8), nt);
        }
    }
    p.ny = (p.nsel + B - 1) / B;
    const long ntiles = p.nb * p.ny;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_extract8<L, B, PF>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid));
    return launch_tile_kernel(k_extract8<L, B, PF>, grid, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L, 8), ntiles);
}
template <int L> static cudaError_t go_extract32(const ExtractParams& p0, cudaStream_t s)
{
    constexpr int B = 4096 / L;
    typedef TileFFT<L, B, -1, false, false, 32> ENG;
    static_assert(ENG::T == 128, "4096-point tiles on 128 threads");
    ExtractParams p = p0;
    {
        long nt = 0;
        if (use_packed(p, B, &nt)) {
            unsigned g = 1;
            FDC_CHECK(persistent_grid(k_extract32_packed<L, B>, ENG::T, tile_smem_bytes<ENG>(), nt, 1, &g, tuning().ctas_ext));
            return launch_tile_kernel(k_extract32_packed<L, B>, g, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L, 32), nt);
        }
    }
    p.ny = (p.nsel + B - 1) / B;
    const long ntiles = p.nb * p.ny;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_extract32<L, B>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid, tuning().ctas_ext));
    return launch_tile_kernel(k_extract32<L, B>, grid, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L, 32), ntiles);
}
template <int L> static cudaError_t go_extract_pf(const ExtractParams& p, cudaStream_t s)
{
    if constexpr (L == 512 || L == 1024) {
        if (tuning().extract_e32) return go_extract32<L>(p, s);
    }
    if constexpr (L >= 64 && L <= 2048) {
        /* short slices have many signals per tile and little work per signal: twice the warps wins there (measured: l = 64
         * 1.5x faster; l = 128 11 % slower on cfg5 since the next tile is prefetched into L2; l >= 256 10-20 % slower);
         * FDC_EXTRACT_E8 = 0 / 1 forces either engine */
        const int e8 = tuning().extract_e8;
        if (e8 > 0 || (e8 < 0 && L <= 64)) return (tuning().prefetch & 2) ? go_extract8<L, true>(p, s) : go_extract8<L, false>(p, s);
    }
    if constexpr (can_prefetch(TileFFT<L, tile_batch(L), -1, false, false>::T)) {
        if (tuning().prefetch & 2) return go_extract<L, true>(p, s);
    }
    return go_extract<L, false>(p, s);
}
template <int L> static cudaError_t go_jobs(const JobParams& p, cudaStream_t s)
{
    constexpr int B = tile_batch(L);
    typedef TileFFT<L, B, -1, false, false> ENG;
    const long ntiles = ((long)p.njobs + B - 1) / B;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_jobs<L, B>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid));
    return launch_tile_kernel(k_jobs<L, B>, grid, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L), ntiles);
}
bool tile_len_supported(int L) { return L >= 2 && L <= 16384 && (L & (L - 1)) == 0; }

#define FDC_FOR_TILE_LENGTHS(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)

cudaError_t launch_extract(const ExtractParams& p, int l, cudaStream_t s)
{
    if (p.nb <= 0 || p.nsel <= 0) return cudaSuccess;
    switch (l) {
#define X(LL) case LL: return go_extract_pf<LL>(p, s);
        FDC_FOR_TILE_LENGTHS(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_jobs(const JobParams& p, int l, cudaStream_t s)
{
    if (p.njobs <= 0) return cudaSuccess;
    switch (l) {
#define X(LL) case LL: return go_jobs<LL>(p, s);
        FDC_FOR_TILE_LENGTHS(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace fdc
