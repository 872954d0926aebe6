/* fdc_k_extract.cu -- K2: batched channel extraction (bin gather, table multiply, half swap, backward FFT,
 * overlap discard, gain), all channel-blocks of one slice length in one launch.  Replaces
 * vector_cut_vxx + phase_shifting_windowing_vcc (lib/vector_cut_vxx_impl.cc:59-72,
 * lib/phase_shifting_windowing_vcc_impl.cc:72-86) with the third-party inverse fft_vcc / multiply_const stages
 * between and behind them, and process_channel of the activity-gated blocks. */
#include "fdc_kcommon.cuh"

namespace fdc {

template <int L, int B, bool PF>
__global__ void __launch_bounds__((TileFFT<L, B, -1, false, false>::T), min_ctas(TileFFT<L, B, -1, false, false>::T, PF))
k_extract(const ExtractParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false>, PF>(ExtractTiles<L, B>{p}, tw, ntiles);
}
template <int L, int B>
__global__ void __launch_bounds__((TileFFT<L, B, -1, false, false>::T), min_ctas(TileFFT<L, B, -1, false, false>::T, false))
k_jobs(const JobParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, -1, false, false>, false>(JobTiles<L, B>{p}, tw, ntiles);
}

template <int L, bool PF> static cudaError_t go_extract(const ExtractParams& p0, cudaStream_t s)
{
    constexpr int B = tile_batch(L);
    typedef TileFFT<L, B, -1, false, false> ENG;
    ExtractParams p = p0;
    p.ny = (p.nsel + B - 1) / B;
    const long ntiles = p.nb * p.ny;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_extract<L, B, PF>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid));
    k_extract<L, B, PF><<<grid, ENG::T, tile_smem_bytes<ENG>(), s>>>(p, twiddle_table(L), ntiles);
    count_launch();
    return cudaGetLastError();
}
template <int L> static cudaError_t go_extract_pf(const ExtractParams& p, cudaStream_t s)
{
    if constexpr (can_prefetch(TileFFT<L, tile_batch(L), -1, false, false>::T)) {
        if (tuning().prefetch) return go_extract<L, true>(p, s);
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
    k_jobs<L, B><<<grid, ENG::T, tile_smem_bytes<ENG>(), s>>>(p, twiddle_table(L), ntiles);
    count_launch();
    return cudaGetLastError();
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
