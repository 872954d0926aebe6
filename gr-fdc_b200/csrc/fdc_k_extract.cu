/* fdc_k_extract.cu -- K2: batched channel extraction (bin gather, table multiply, half swap, backward FFT,
 * overlap discard, gain), all channel-blocks of one slice length in one launch.  Replaces
 * vector_cut_vxx + phase_shifting_windowing_vcc (lib/vector_cut_vxx_impl.cc:59-72,
 * lib/phase_shifting_windowing_vcc_impl.cc:72-86) with the third-party inverse fft_vcc / multiply_const stages
 * between and behind them, and process_channel of the activity-gated blocks. */
#include "fdc_kcommon.cuh"

namespace fdc {

template <int L, int B>
__global__ void __launch_bounds__((TileFFT<L, B, -1, false, false>::T), min_ctas(TileFFT<L, B, -1, false, false>::T))
k_extract(const ExtractParams p, const float2* __restrict__ tw)
{
    typedef TileFFT<L, B, -1, false, false> ENG;
    ExtractLoader<L, B> ld{p, (int)blockIdx.x, (int)blockIdx.y};
    ExtractStorer<L, B> st{p, (int)blockIdx.x, (int)blockIdx.y};
    tile_fft_run<ENG>(reinterpret_cast<float2*>(fdc_smem_raw), tw, ld, st);
}
template <int L, int B>
__global__ void __launch_bounds__((TileFFT<L, B, -1, false, false>::T), min_ctas(TileFFT<L, B, -1, false, false>::T))
k_jobs(const JobParams p, const float2* __restrict__ tw)
{
    typedef TileFFT<L, B, -1, false, false> ENG;
    JobLoader<L, B> ld{p, (int)blockIdx.x};
    JobStorer<L, B> st{p, (int)blockIdx.x};
    tile_fft_run<ENG>(reinterpret_cast<float2*>(fdc_smem_raw), tw, ld, st);
}

template <int L> static cudaError_t go_extract(const ExtractParams& p, int nsel, cudaStream_t s)
{
    constexpr int B = tile_batch(L);
    typedef TileFFT<L, B, -1, false, false> ENG;
    FDC_CHECK(set_smem(k_extract<L, B>, ENG::SMEM_BYTES));
    const unsigned gx = (unsigned)((p.nb + B - 1) / B);
    for (int y0 = 0; y0 < nsel; y0 += 65535) {
        ExtractParams q = p; q.sel = p.sel + y0;
        const int ny = nsel - y0 < 65535 ? nsel - y0 : 65535;
        k_extract<L, B><<<dim3(gx, (unsigned)ny), ENG::T, ENG::SMEM_BYTES, s>>>(q, twiddle_table(L));
        count_launch();
    }
    return cudaGetLastError();
}
template <int L> static cudaError_t go_jobs(const JobParams& p, cudaStream_t s)
{
    constexpr int B = tile_batch(L);
    typedef TileFFT<L, B, -1, false, false> ENG;
    FDC_CHECK(set_smem(k_jobs<L, B>, ENG::SMEM_BYTES));
    const unsigned gx = (unsigned)((p.njobs + B - 1) / B);
    k_jobs<L, B><<<gx, ENG::T, ENG::SMEM_BYTES, s>>>(p, twiddle_table(L));
    count_launch();
    return cudaGetLastError();
}
bool tile_len_supported(int L) { return L >= 2 && L <= 16384 && (L & (L - 1)) == 0; }

#define FDC_FOR_TILE_LENGTHS(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)

cudaError_t launch_extract(const ExtractParams& p, int l, int nsel, cudaStream_t s)
{
    if (p.nb <= 0 || nsel <= 0) return cudaSuccess;
    switch (l) {
#define X(LL) case LL: return go_extract<LL>(p, nsel, s);
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
