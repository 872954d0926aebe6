/* fdc_k_fwd.cu -- K1: overlap-save staging + forward FFT + fft-shift + 1/N, one kernel (N <= 16384)
 * or the two kernels of the four-step scheme (N >= 32768).  Replaces FDC.overlap_save
 * (lib/overlap_save_impl.cc:62-81) and the third-party fft_vcc / multiply_const stages the hier block
 * puts behind it (python/FrequencyDomainChannelizer.py:202-216).
 * All kernels are persistent (grid = resident CTAs) and prefetch the next tile's samples into registers
 * while the current tile is transformed. */
#include "fdc_kcommon.cuh"

namespace fdc {

template <int N, int B, bool PF>
__global__ void __launch_bounds__((TileFFT<N, B, 1, false, false>::T), min_ctas(TileFFT<N, B, 1, false, false>::T, PF, N))
k_fwd_small(const FwdParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<N, B, 1, false, false>, PF>(FwdTiles<N, B>{p}, tw, ntiles);
}

/* E = 32 (lengths 512 / 1024: two passes instead of three) runs without prefetch on 128 threads, four CTAs per SM */
template <int N1, int N2, int B, bool PF, int E>
__global__ void __launch_bounds__((TileFFT<N1, B, 1, true, true, E>::T), E == 32 ? 4 : min_ctas(TileFFT<N1, B, 1, true, true, E>::T, PF, N1))
k_fwd_cols(const BigParams p, const float2* __restrict__ tw, long ntiles)
{
    typedef TileFFT<N1, B, 1, true, true, E> ENG;
    /* the launcher rounds the grid to a multiple of the column tiles per block: this CTA keeps column tile blockIdx % (N2/B) */
    float2* tws = reinterpret_cast<float2*>(fdc_smem_raw) + ENG::SMEM_ELEMS + tw_smem_elems(ENG::L, ENG::E);
    ENG::template last_pass_init<ColTwiddles<N1, N2, B> >((int)threadIdx.x, tws, p.tw4, (int)(blockIdx.x % (N2 / B)));
    tile_kernel_body<ENG, PF>(ColTiles<N1, N2, B>{p, tws}, tw, ntiles);
}
template <int N1, int N2, int B, bool PF, int E>
__global__ void __launch_bounds__((TileFFT<N2, B, 1, false, true, E>::T), E == 32 ? 4 : min_ctas(TileFFT<N2, B, 1, false, true, E>::T, PF, N2))
k_fwd_rows(const BigParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<N2, B, 1, false, true, E>, PF>(RowTiles<N1, N2, B>{p}, tw, ntiles);
}

/* 32 points per thread where that saves a pass: 512 / 1024 (two instead of three), 8192 / 16384 (three instead of four) */
template <int N, int B>
__global__ void __launch_bounds__(N * B / 32, N * B / 32 <= 128 ? 4 : (N * B / 32 <= 256 ? 2 : 1))
k_fwd_small32(const FwdParams p, const float2* __restrict__ tw, long ntiles)
{
    tile_kernel_body<TileFFT<N, B, 1, false, false, 32>, false>(FwdTiles<N, B>{p}, tw, ntiles);
}
template <int N> static cudaError_t go_small32(const FwdParams& p, cudaStream_t s)
{
    constexpr int B = tile_batch(N);
    typedef TileFFT<N, B, 1, false, false, 32> ENG;
    const long ntiles = (p.nblocks + B - 1) / B;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_fwd_small32<N, B>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid, tuning().ctas_fwd));
    return launch_tile_kernel(k_fwd_small32<N, B>, grid, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(N, 32), ntiles);
}
template <int N, bool PF> static cudaError_t go_small(const FwdParams& p, cudaStream_t s)
{
    constexpr int B = tile_batch(N);
    typedef TileFFT<N, B, 1, false, false> ENG;
    const long ntiles = (p.nblocks + B - 1) / B;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_fwd_small<N, B, PF>, ENG::T, tile_smem_bytes<ENG>(), ntiles, 1, &grid));
    return launch_tile_kernel(k_fwd_small<N, B, PF>, grid, ENG::T, tile_smem_bytes<ENG>(), s, p, twiddle_table(N), ntiles);
}
template <int N> static cudaError_t go_small_pf(const FwdParams& p, cudaStream_t s)
{
    if constexpr (can_prefetch(TileFFT<N, tile_batch(N), 1, false, false>::T)) {
        if (tuning().prefetch & 1) return go_small<N, true>(p, s);
    }
    return go_small<N, false>(p, s);
}
bool fwd_small_supported(int N) { return N >= 16 && N <= 16384 && (N & (N - 1)) == 0; }
cudaError_t launch_fwd_small(const FwdParams& p, cudaStream_t s)
{
    if (p.nblocks <= 0) return cudaSuccess;
    switch (p.N) {
    case 16: return go_small_pf<16>(p, s);
    case 32: return go_small_pf<32>(p, s);
    case 64: return go_small_pf<64>(p, s);
    case 128: return go_small_pf<128>(p, s);
    case 256: return go_small_pf<256>(p, s);
    case 512: return tuning().fwd_e32 > 1 ? go_small32<512>(p, s) : go_small_pf<512>(p, s);     /* measured: no gain at 1024 */
    case 1024: return tuning().fwd_e32 > 1 ? go_small32<1024>(p, s) : go_small_pf<1024>(p, s);
    case 2048: return go_small_pf<2048>(p, s);
    case 4096: return go_small_pf<4096>(p, s);
    case 8192: return tuning().fwd_e32 ? go_small32<8192>(p, s) : go_small_pf<8192>(p, s);
    case 16384: return tuning().fwd_e32 ? go_small32<16384>(p, s) : go_small_pf<16384>(p, s);
    }
    return cudaErrorInvalidValue;
}

constexpr int big_points(int L, bool e32) { return (e32 && (L == 512 || L == 1024)) ? 32 : 16; }     /* points per thread of a column / row transform */
template <int N1, int N2, bool PF, bool E32 = true> static cudaError_t go_big(const BigParams& p, cudaStream_t s)
{
    if constexpr (E32 && (N1 >= 512 || N2 >= 512)) {
        /* measured on cfg5 (512 x 512): 16 points + prefetch 60 Gsample/s, 32 points without prefetch 52: these kernels
         * are bandwidth bound, the prefetch counts for more than the saved pass.  FDC_FWD_E32=2 selects the 32-point form. */
        if (tuning().fwd_e32 < 2) return go_big<N1, N2, PF, false>(p, s);
    }
    /* 16 adjacent columns / rows (128-byte lines) while the tile fits 256 threads, narrower tiles above that */
    constexpr int BC = big_tile_batch(N1), BR = big_tile_batch(N2);
    constexpr int EC = big_points(N1, E32), ER = big_points(N2, E32);
    constexpr bool PFC = PF && EC == 16, PFR = PF && ER == 16;
    typedef TileFFT<N1, BC, 1, true, true, EC> CE;
    typedef TileFFT<N2, BR, 1, false, true, ER> RE;
    constexpr size_t csmem = tile_smem_bytes<CE>() + sizeof(float2) * (size_t)N1 * BC;   /* exchange buffer + this CTA's four-step twiddles */
    const long ctiles = p.nblocks * (N2 / BC), rtiles = p.nblocks * (N1 / BR);
    unsigned gc = 1, gr = 1;
    /* a column CTA keeps its column tile and with it its slice of the four-step twiddle table */
    FDC_CHECK(persistent_grid(k_fwd_cols<N1, N2, BC, PFC, EC>, CE::T, csmem, ctiles, N2 / BC, &gc, tuning().ctas_fwd));
    FDC_CHECK(persistent_grid(k_fwd_rows<N1, N2, BR, PFR, ER>, RE::T, tile_smem_bytes<RE>(), rtiles, 1, &gr, tuning().ctas_fwd));
    FDC_CHECK(launch_tile_kernel(k_fwd_cols<N1, N2, BC, PFC, EC>, gc, CE::T, csmem, s, p, twiddle_table(N1, EC), ctiles));
    return launch_tile_kernel(k_fwd_rows<N1, N2, BR, PFR, ER>, gr, RE::T, tile_smem_bytes<RE>(), s, p, twiddle_table(N2, ER), rtiles);
}
bool fwd_big_supported(int N, int* N1, int* N2)
{
    int a = 0, b = 0;
    switch (N) {
    case 4096: a = 64; b = 64; break;
    case 8192: a = 64; b = 128; break;
    case 16384: a = 128; b = 128; break;
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
    const bool pf = (tuning().prefetch & 1) != 0;
    switch (N) {
#define X(NN, A, BB) case NN: return pf ? go_big<A, BB, true>(p, s) : go_big<A, BB, false>(p, s);
    X(4096, 64, 64) X(8192, 64, 128) X(16384, 128, 128) X(32768, 128, 256) X(65536, 256, 256) X(131072, 256, 512) X(262144, 512, 512) X(524288, 512, 1024) X(1048576, 1024, 1024)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace fdc
