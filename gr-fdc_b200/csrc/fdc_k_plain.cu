/* fdc_k_plain.cu -- batched fft_vcc stage (rectangular window, optional shift) for flowgraphs that keep the
 * third-party FFT blocks separate (python/FrequencyDomainChannelizer.py:206,228). */
#include "fdc_kcommon.cuh"

namespace fdc {

template <int L, int B, int DIR>
__global__ void __launch_bounds__((TileFFT<L, B, DIR, false, false>::T), min_ctas(TileFFT<L, B, DIR, false, false>::T))
k_plain(const PlainParams p, const float2* __restrict__ tw)
{
    typedef TileFFT<L, B, DIR, false, false> ENG;
    PlainLoader<L, B, DIR> ld{p, (int)blockIdx.x};
    PlainStorer<L, B, DIR> st{p, (int)blockIdx.x};
    tile_fft_run<ENG>(reinterpret_cast<float2*>(fdc_smem_raw), tw, ld, st);
}
template <int L, int DIR> static cudaError_t go_plain(const PlainParams& p, cudaStream_t s)
{
    constexpr int B = tile_batch(L);
    typedef TileFFT<L, B, DIR, false, false> ENG;
    FDC_CHECK(set_smem(k_plain<L, B, DIR>, ENG::SMEM_BYTES));
    const long tiles = (p.nvec + B - 1) / B;
    for (long t0 = 0; t0 < tiles; t0 += 1 << 30) {
        PlainParams q = p; q.in = p.in + t0 * B * L; q.out = p.out + t0 * B * L; q.nvec = p.nvec - t0 * B;
        const long nt = tiles - t0 < (1 << 30) ? tiles - t0 : (1 << 30);
        k_plain<L, B, DIR><<<(unsigned)nt, ENG::T, ENG::SMEM_BYTES, s>>>(q, twiddle_table(L));
        count_launch();
    }
    return cudaGetLastError();
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
