/* fdc_k_fwd.cu -- K1: overlap-save staging + forward FFT + fft-shift + 1/N, one kernel (N <= 16384)
 * or the two kernels of the four-step scheme (N >= 32768).  Replaces FDC.overlap_save
 * (lib/overlap_save_impl.cc:62-81) and the third-party fft_vcc / multiply_const stages the hier block
 * puts behind it (python/FrequencyDomainChannelizer.py:202-216). */
#include "fdc_kcommon.cuh"

namespace fdc {

template <int N, int B>
__global__ void __launch_bounds__((TileFFT<N, B, 1, false, false>::T), min_ctas(TileFFT<N, B, 1, false, false>::T))
k_fwd_small(const FwdParams p, const float2* __restrict__ tw)
{
    typedef TileFFT<N, B, 1, false, false> ENG;
    FwdLoader<N, B> ld{p, (int)blockIdx.x};
    FwdStorer<N, B> st{p, (int)blockIdx.x};
    tile_fft_run<ENG>(reinterpret_cast<float2*>(fdc_smem_raw), tw, ld, st);
}

template <int N1, int N2, int B>
__global__ void __launch_bounds__((TileFFT<N1, B, 1, true, true>::T), min_ctas(TileFFT<N1, B, 1, true, true>::T))
k_fwd_cols(const BigParams p, const float2* __restrict__ tw)
{
    typedef TileFFT<N1, B, 1, true, true> ENG;
    ColLoader<N1, N2, B> ld{p, (int)blockIdx.x, (long)blockIdx.y};
    ColStorer<N1, N2, B> st{p, (int)blockIdx.x, (long)blockIdx.y};
    tile_fft_run<ENG>(reinterpret_cast<float2*>(fdc_smem_raw), tw, ld, st);
}
template <int N1, int N2, int B>
__global__ void __launch_bounds__((TileFFT<N2, B, 1, false, true>::T), min_ctas(TileFFT<N2, B, 1, false, true>::T))
k_fwd_rows(const BigParams p, const float2* __restrict__ tw)
{
    typedef TileFFT<N2, B, 1, false, true> ENG;
    RowLoader<N1, N2, B> ld{p, (int)blockIdx.x, (long)blockIdx.y};
    RowStorer<N1, N2, B> st{p, (int)blockIdx.x, (long)blockIdx.y};
    tile_fft_run<ENG>(reinterpret_cast<float2*>(fdc_smem_raw), tw, ld, st);
}

template <int N> static cudaError_t go_small(const FwdParams& p, cudaStream_t s)
{
    constexpr int B = tile_batch(N);
    typedef TileFFT<N, B, 1, false, false> ENG;
    FDC_CHECK(set_smem(k_fwd_small<N, B>, ENG::SMEM_BYTES));
    const unsigned grid = (unsigned)((p.nblocks + B - 1) / B);
    k_fwd_small<N, B><<<grid, ENG::T, ENG::SMEM_BYTES, s>>>(p, twiddle_table(N));
    count_launch();
    return cudaGetLastError();
}
bool fwd_small_supported(int N) { return N >= 16 && N <= 16384 && (N & (N - 1)) == 0; }
cudaError_t launch_fwd_small(const FwdParams& p, cudaStream_t s)
{
    if (p.nblocks <= 0) return cudaSuccess;
    switch (p.N) {
    case 16: return go_small<16>(p, s);
    case 32: return go_small<32>(p, s);
    case 64: return go_small<64>(p, s);
    case 128: return go_small<128>(p, s);
    case 256: return go_small<256>(p, s);
    case 512: return go_small<512>(p, s);
    case 1024: return go_small<1024>(p, s);
    case 2048: return go_small<2048>(p, s);
    case 4096: return go_small<4096>(p, s);
    case 8192: return go_small<8192>(p, s);
    case 16384: return go_small<16384>(p, s);
    }
    return cudaErrorInvalidValue;
}

template <int N1, int N2> static cudaError_t go_big(const BigParams& p, cudaStream_t s)
{
    constexpr int B = 16;
    typedef TileFFT<N1, B, 1, true, true> CE;
    typedef TileFFT<N2, B, 1, false, true> RE;
    FDC_CHECK(set_smem(k_fwd_cols<N1, N2, B>, CE::SMEM_BYTES));
    FDC_CHECK(set_smem(k_fwd_rows<N1, N2, B>, RE::SMEM_BYTES));
    /* gridDim.y is limited to 65535 blocks per launch */
    for (long b0 = 0; b0 < p.nblocks; b0 += 32768) {
        BigParams q = p;
        const long nb = p.nblocks - b0 < 32768 ? p.nblocks - b0 : 32768;
        q.in = p.in + b0 * p.hop; q.mid = p.mid + b0 * (long)N1 * N2; q.spec = p.spec + b0 * (long)N1 * N2; q.nblocks = nb;
        if (b0 > 0) { q.hist = p.in + b0 * p.hop - p.ovl; }
        k_fwd_cols<N1, N2, B><<<dim3(N2 / B, (unsigned)nb), CE::T, CE::SMEM_BYTES, s>>>(q, twiddle_table(N1));
        k_fwd_rows<N1, N2, B><<<dim3(N1 / B, (unsigned)nb), RE::T, RE::SMEM_BYTES, s>>>(q, twiddle_table(N2));
        count_launch(2);
    }
    return cudaGetLastError();
}
bool fwd_big_supported(int N, int* N1, int* N2)
{
    int a = 0, b = 0;
    switch (N) {
    case 32768: a = 128; b = 256; break;
    case 65536: a = 256; b = 256; break;
    case 131072: a = 256; b = 512; break;
    case 262144: a = 512; b = 512; break;
    case 524288: a = 512; b = 1024; break;
    case 1048576: a = 1024; b = 1024; break;
    default: return false;
    }
    if (N1) *N1 = a;
    if (N2) *N2 = b;
    return true;
}
cudaError_t launch_fwd_big(const BigParams& p, int N, cudaStream_t s)
{
    if (p.nblocks <= 0) return cudaSuccess;
    switch (N) {
    case 32768: return go_big<128, 256>(p, s);
    case 65536: return go_big<256, 256>(p, s);
    case 131072: return go_big<256, 512>(p, s);
    case 262144: return go_big<512, 512>(p, s);
    case 524288: return go_big<512, 1024>(p, s);
    case 1048576: return go_big<1024, 1024>(p, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fdc
