/* fdc_k_plain.cu -- batched fft_vcc stage (rectangular window, optional shift) for flowgraphs that keep the
 * third-party FFT blocks separate (python/FrequencyDomainChannelizer.py:206,228). */
#include "fdc_kcommon.cuh"

namespace fdc {

template <int L, int B, int DIR>
__global__ void __launch_bounds__((TileFFT<L, B, DIR, false, false>::T), min_ctas(TileFFT<L, B, DIR, false, false>::T, false, L))
k_plain(const PlainParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<L, B, DIR, false, false>, false>(PlainTiles<L, B, DIR>{p}, tw, ntiles);
}
template <int L, int DIR> static cudaError_t go_plain(const PlainParams& p, cudaStream_t s)
{
    constexpr int B = tile_batch(L);
    typedef TileFFT<L, B, DIR, false, false> ENG;
    const long ntiles = (p.nvec + B - 1) / B;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_plain<L, B, DIR>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid));
    return launch_tile_kernel(k_plain<L, B, DIR>, grid, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(L), ntiles);
}
cudaError_t launch_plain_fft(const PlainParams& p, int L, int forward, cudaStream_t s)
{
    if (p.nvec <= 0) return cudaSuccess;
    switch (L) {
#define X(LL) case LL: return forward ? go_plain<LL, 1>(p, s) : go_plain<LL, -1>(p, s);
        X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace fdc
