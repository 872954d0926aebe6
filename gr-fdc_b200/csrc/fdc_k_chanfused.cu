/* fdc_k_chanfused.cu -- K1 + K2 in ONE kernel for FFT lengths whose block fits a CTA's shared memory (N <= 16384):
 * overlap-save staging, forward FFT, fft-shift and 1/N, and then -- out of shared memory, the spectrum never leaves the SM --
 * every channel's bin cut, table multiply, half swap, backward FFT, overlap discard and gain.  Replaces the whole chain
 * overlap_save -> fft_vcc -> multiply_const -> {vector_cut -> phase_shifting_windowing_vcc -> fft_vcc -> vector_cut ->
 * multiply_const} x channels of the hier block (python/FrequencyDomainChannelizer.py:201-231) with one launch per chunk.
 *
 * Why: the kernels are bound by the SMs' interface to L2 (about 29 B per clock and SM for loads and stores together,
 * DESIGN.md 4).  The two-kernel form sends the spectrum through it twice (8 N / hop B per input sample out, and the slices,
 * 1.2 x that with 2 x oversampled channels, back in); fused, only the samples come in and only the channel outputs go out:
 * cfg2 (N = 8192, 64 channels of 256 bins) 50.1 -> 26.7 B per input sample.
 *
 * A CTA walks forward tiles (BF blocks of N points, the tile of k_fwd_small) persistently.  Per tile:
 *   1. fetch the samples (two-segment loads for the blocks that reach into the history), all forward passes but the last
 *      through the exchange region R1;
 *   2. last forward pass: operands out of R1 into registers, barrier, results -- shifted and scaled -- back into R1, which is
 *      now the spectrum S[BF][N] of the tile;
 *   3. the signals (block of the tile, channel) are walked in extract tiles of BX signals: slices out of S, table from
 *      global memory (L1 resident), backward transform through the exchange region R2, rows to the outputs (or sinks).
 * One slice length per launch (every channel of the context has the same l); contexts with several slice lengths, lengths
 * outside the table below, or a caller who wants the spectrum (debug port, activity blocks) take the two-kernel path. */
#include "fdc_kcommon.cuh"
#include <cstdio>

namespace fdc {

/* phases [PH, END) of a tile transform with CTA barriers in between (tile_fft_from runs to the end) */
template <class ENG, int PH, int END, bool TWS, class Storer>
__device__ __forceinline__ void tile_fft_range(float2* v, float2* smem, const float2* tw, const Storer& st)
{
    if constexpr (PH < END) {
        ENG::template phase<PH, TWS>(threadIdx.x, v, smem, tw, st);
        __syncthreads();
        tile_fft_range<ENG, PH + 1, END, TWS, Storer>(v, smem, tw, st);
    }
}

/* last forward pass -> spectrum in shared memory: fft_vcc shift (out[0:N/2] = Y[N/2:N], ...) and 1/N as FwdStorer does */
template <int N> struct SmemSpecStorer {
    typedef float2* Ctx;
    float2* s; float scale;
    __device__ __forceinline__ Ctx begin(int batch, int o) const { return s + (batch * N + o); }
    template <int R, int NS> __device__ __forceinline__ void put(const Ctx& row, int t, float2 v) const { row[(t ^ (R / 2)) * NS] = cscale(v, scale); }
};

/* signal s of a forward tile = (block s / nch of the tile, channel s % nch) */
struct FusedSignal {
    int blk, chan; bool valid;
    __device__ __forceinline__ FusedSignal(int s, int nsig, int nch)
    {
        valid = s < nsig;
        if (!valid) s = nsig - 1;                           /* padding signals recompute the last one, nothing is stored */
        blk = s / nch; chan = s - blk * nch;
    }
};
template <int N, int L, int BX> struct FusedExtractLoader {
    struct Ctx { const float2* x; const float2* w; };
    static constexpr bool HAS_FINISH = true;
    const ExtractParams& p; const float2* spec; long blk0; int xt, nsig;        /* blk0: first block of the forward tile inside the chunk */
    __device__ __forceinline__ Ctx begin(int batch, int j) const
    {
        Ctx c;
        const FusedSignal g(xt * BX + batch, nsig, p.nsel);
        const ChanDev& ch = p.chans[g.chan];
        const unsigned bphase = phase_mod(p, (unsigned)p.glob_phase0 + (unsigned)(blk0 + g.blk));
        const unsigned phase = phase_mod(p, bphase * (unsigned)ch.shift);
        c.x = spec + (g.blk * N + ch.f + j);
        c.w = p.tables + (ch.tab_off + (long)phase * L + j);
        return c;
    }
    /* fft_vcc inverse+shift: FFT input n takes bin (n + l/2) mod l */
    template <int R, int STRIDE> __device__ __forceinline__ float2 fetch(const Ctx& c, int t) const { return c.x[((t + R / 2) % R) * STRIDE]; }
    template <int R, int STRIDE> __device__ __forceinline__ float2 finish(const Ctx& c, int t, float2 raw) const
    {
        return cmul(raw, fdc_ldg(c.w + ((t + R / 2) % R) * STRIDE));
    }
};
template <int N, int L, int BX> struct FusedExtractStorer {
    struct Ctx { float2* dst; int skip; float gain; };
    const ExtractParams& p; long blk0; int xt, nsig;
    __device__ __forceinline__ Ctx begin(int batch, int o) const
    {
        Ctx c; c.dst = p.out; c.skip = 1 << 30; c.gain = 0.f;
        const FusedSignal g(xt * BX + batch, nsig, p.nsel);
        if (!g.valid) return c;
        const ChanDev& ch = p.chans[g.chan];
        c.skip = L - ch.lout - o; c.gain = ch.gain;
        c.dst = chan_out_row(p, ch, blk0 + g.blk) - (L - ch.lout) + o;
        return c;
    }
    static constexpr bool HAS_VARIANT = true;
    __device__ __forceinline__ bool variant(const Ctx& c) const { return c.gain == 1.0f; }
    template <int R, int NS, bool UNIT> __device__ __forceinline__ void put(const Ctx& c, int t, float2 v) const
    {
        if (t * NS >= c.skip) c.dst[t * NS] = UNIT ? v : cscale(v, c.gain);
    }
};

template <int N, int BF, int EF, int L, int EX> struct ChanFused {
    typedef TileFFT<N, BF, 1, false, false, EF> FE;
    static constexpr int T = FE::T;
    static constexpr int BX = T * EX / L;                    /* signals per extract tile */
    typedef TileFFT<L, BX, -1, false, false, EX> XE;
    static_assert(XE::T == T && BX >= 1, "forward and extract tiles use the same CTA");
    static_assert(FE::NP >= 2 && XE::NP >= 2, "both transforms exchange through shared memory");
    static constexpr int R1_ELEMS = (FE::SMEM_ELEMS > BF * N ? FE::SMEM_ELEMS : BF * N);     /* forward exchange, then the spectrum */
    static constexpr int R2_ELEMS = XE::SMEM_ELEMS;
    static constexpr size_t SMEM_BYTES = sizeof(float2) * (size_t)(R1_ELEMS + R2_ELEMS + FE::TWSIZE + XE::TWSIZE);
    static constexpr int CTAS = (T * (EF > EX ? EF : EX) * 2 + 48 * T) <= 65536 / 2 ? 2 : 1;    /* register budget: 2 CTAs when 2 x T x (data + ~48) fits */
};

template <int N, int BF, int EF, int L, int EX>
__global__ void __launch_bounds__((ChanFused<N, BF, EF, L, EX>::T), (ChanFused<N, BF, EF, L, EX>::CTAS))
k_chan_fused(const FwdParams fp, const ExtractParams xp, const float2* __restrict__ twf_g, const float2* __restrict__ twx_g, long ntiles)
{
    typedef ChanFused<N, BF, EF, L, EX> K;
    typedef typename K::FE FE; typedef typename K::XE XE;
    constexpr int BX = K::BX;
    float2* R1 = reinterpret_cast<float2*>(fdc_smem_raw);
    float2* R2 = R1 + K::R1_ELEMS;
    float2* twf = R2 + K::R2_ELEMS;
    float2* twx = twf + FE::TWSIZE;
    const int tid = (int)threadIdx.x;
    cudaTriggerProgrammaticLaunchCompletion();
    for (int i = tid; i < FE::TWSIZE; i += K::T) twf[i] = twf_g[i];
    for (int i = tid; i < XE::TWSIZE; i += K::T) twx[i] = twx_g[i];
    __syncthreads();
    cudaGridDependencySynchronize();
    const FwdTiles<N, BF> ftiles{fp};
    const SmemSpecStorer<N> to_spec{R1, fp.scale};
#if defined(FDC_FUSED_PROF)
    long long t_f0 = 0, t_f1 = 0, t_x = 0, t_c = clock64(); long nt = 0;
#define FDC_TK(acc) do { const long long n_ = clock64(); acc += n_ - t_c; t_c = n_; } while (0)
#else
#define FDC_TK(acc) do { } while (0)
#endif
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TilePos pos; pos.inner = 0; pos.outer = (int)tile;
        /* ---- forward transform of the tile's blocks; the spectrum ends up in R1 ---- */
        {
            float2 v[FE::E];
            FE::fetch(tid, v, ftiles.loader(pos));
            if (tile + gridDim.x < ntiles) { TilePos np; np.inner = 0; np.outer = (int)(tile + gridDim.x); ftiles.prefetch_l2(np, tid); }
            FE::template phase<0, true>(tid, v, R1, twf, to_spec);          /* waits for the samples */
            __syncthreads();
            FDC_TK(t_f0);
            tile_fft_range<FE, 1, FE::NPH - 1, true>(v, R1, twf, to_spec);
            FE::template read_smem<FE::NP - 1>(tid, v, R1);
            __syncthreads();                                 /* every operand of the last pass is in registers: R1 becomes the spectrum */
            FE::template twiddle_bfly<FE::NP - 1, true>(tid, v, twf);
            FE::template store_global<FE::NP - 1>(tid, v, to_spec);
            __syncthreads();
            FDC_TK(t_f1);
        }
        /* ---- all channels of the tile's blocks ---- */
        const long blk0 = tile * BF;
        const long nblk = fp.nblocks - blk0 < BF ? fp.nblocks - blk0 : BF;
        const int nsig = (int)nblk * xp.nsel;
        const int nxt = (nsig + BX - 1) / BX;
        for (int xt = 0; xt < nxt; xt++) {
            float2 x[XE::E];
            const FusedExtractLoader<N, L, BX> ld{xp, R1, blk0, xt, nsig};
            XE::fetch(tid, x, ld);
            XE::finish(tid, x, ld);
            tile_fft_from<XE, 0, true>(x, R2, twx, FusedExtractStorer<N, L, BX>{xp, blk0, xt, nsig});
            __syncthreads();                                 /* R2 is reused by the next extract tile, R1 by the next forward tile */
        }
        FDC_TK(t_x);
#if defined(FDC_FUSED_PROF)
        nt++;
#endif
    }
#if defined(FDC_FUSED_PROF)
    if (blockIdx.x == 7 && tid == 0 && nt > 0) printf("fused tile: load + first pass %lld, rest of forward %lld, extract %lld cycles (%ld tiles)\n", t_f0 / nt, t_f1 / nt, t_x / nt, nt);
#endif
#undef FDC_TK
}

/* ---- the same with the two halves of the work on two halves of the CTA (FDC_FUSE_SMALL=2) --------------------------------
 * The one-group form runs load -> forward passes -> extract tiles -> stores strictly one after the other in a CTA, and only two
 * CTAs fit an SM: nothing overlaps the forward transform's loads or the extract's stores.  Here a CTA has 2 T threads: group F
 * (threads 0 .. T-1) transforms block n + 1 into spectrum buffer S[(n + 1) & 1] while group X (threads T .. 2T-1) takes the
 * channels of block n out of S[n & 1].  Producer / consumer hand-over with named barriers (bar.arrive by the one group, bar.sync
 * by the other, count 2 T): FULL[k] = "S[k] holds a spectrum", EMPTY[k] = "S[k] has been read".  Inside a group the exchange
 * barriers are named barriers of T threads. */
template <int ID, int COUNT> struct GroupSync {
    static __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
};
template <int COUNT> __device__ __forceinline__ void bar_sync_id(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory"); }
template <int COUNT> __device__ __forceinline__ void bar_arrive_id(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(COUNT) : "memory"); }

template <class ENG, int PH, int END, bool TWS, class Sync, class Storer>
__device__ __forceinline__ void group_fft_range(int tid, float2* v, float2* smem, const float2* tw, const Storer& st)
{
    if constexpr (PH < END) {
        ENG::template phase<PH, TWS>(tid, v, smem, tw, st);
        if constexpr (PH + 1 < ENG::NPH) Sync::sync();
        group_fft_range<ENG, PH + 1, END, TWS, Sync, Storer>(tid, v, smem, tw, st);
    }
}

template <int N, int BF, int EF, int L, int EX>
__global__ void __launch_bounds__((2 * ChanFused<N, BF, EF, L, EX>::T), 1)
k_chan_pingpong(const FwdParams fp, const ExtractParams xp, const float2* __restrict__ twf_g, const float2* __restrict__ twx_g, long ntiles)
{
    typedef ChanFused<N, BF, EF, L, EX> K;
    typedef typename K::FE FE; typedef typename K::XE XE;
    constexpr int BX = K::BX, T = K::T;
    enum { BAR_F = 1, BAR_X = 2, BAR_FULL = 3, BAR_EMPTY = 5 };           /* FULL / EMPTY: + buffer index */
    float2* S0 = reinterpret_cast<float2*>(fdc_smem_raw);
    float2* S1 = S0 + K::R1_ELEMS;
    float2* R2 = S1 + K::R1_ELEMS;
    float2* twf = R2 + K::R2_ELEMS;
    float2* twx = twf + FE::TWSIZE;
    const int group = (int)threadIdx.x / T, tid = (int)threadIdx.x % T;
    cudaTriggerProgrammaticLaunchCompletion();
    for (int i = (int)threadIdx.x; i < FE::TWSIZE; i += 2 * T) twf[i] = twf_g[i];
    for (int i = (int)threadIdx.x; i < XE::TWSIZE; i += 2 * T) twx[i] = twx_g[i];
    __syncthreads();
    cudaGridDependencySynchronize();
    long n = 0;
    if (group == 0) {
        /* ---- group F: samples -> spectrum buffers ---- */
        const FwdTiles<N, BF> ftiles{fp};
        typedef GroupSync<BAR_F, T> FS;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n++) {
            const int k = (int)(n & 1);
            float2* S = k ? S1 : S0;
            const SmemSpecStorer<N> to_spec{S, fp.scale};
            TilePos pos; pos.inner = 0; pos.outer = (int)tile;
            float2 v[FE::E];
            FE::fetch(tid, v, ftiles.loader(pos));                       /* in flight while the buffer is still being read */
            if (tile + gridDim.x < ntiles) { TilePos np; np.inner = 0; np.outer = (int)(tile + gridDim.x); ftiles.prefetch_l2(np, tid); }
            if (n >= 2) bar_sync_id<2 * T>(BAR_EMPTY + k);               /* group X has taken the spectrum of tile n - 2 out of S[k] */
            group_fft_range<FE, 0, FE::NPH - 1, true, FS>(tid, v, S, twf, to_spec);
            FE::template read_smem<FE::NP - 1>(tid, v, S);
            FS::sync();                                                  /* every operand of the last pass is in registers: S becomes the spectrum */
            FE::template twiddle_bfly<FE::NP - 1, true>(tid, v, twf);
            FE::template store_global<FE::NP - 1>(tid, v, to_spec);
            bar_arrive_id<2 * T>(BAR_FULL + k);
        }
    } else {
        /* ---- group X: spectrum buffers -> channel outputs ---- */
        typedef GroupSync<BAR_X, T> XS;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n++) {
            const int k = (int)(n & 1);
            const float2* S = k ? S1 : S0;
            bar_sync_id<2 * T>(BAR_FULL + k);
            const long blk0 = tile * BF;
            const long nblk = fp.nblocks - blk0 < BF ? fp.nblocks - blk0 : BF;
            const int nsig = (int)nblk * xp.nsel;
            const int nxt = (nsig + BX - 1) / BX;
            for (int xt = 0; xt < nxt; xt++) {
                float2 x[XE::E];
                const FusedExtractLoader<N, L, BX> ld{xp, S, blk0, xt, nsig};
                XE::fetch(tid, x, ld);
                XE::finish(tid, x, ld);
                group_fft_range<XE, 0, XE::NPH, true, XS>(tid, x, R2, twx, FusedExtractStorer<N, L, BX>{xp, blk0, xt, nsig});
                XS::sync();                                              /* R2 is reused by the next extract tile */
            }
            if (tile + 2 * (long)gridDim.x < ntiles) bar_arrive_id<2 * T>(BAR_EMPTY + k);       /* only when group F will wait for it */
        }
    }
}
template <int N, int BF, int EF, int L, int EX> static cudaError_t go_pingpong(const FwdParams& fp, const ExtractParams& xp, cudaStream_t s)
{
    typedef ChanFused<N, BF, EF, L, EX> K;
    constexpr size_t smem = sizeof(float2) * (size_t)(2 * K::R1_ELEMS + K::R2_ELEMS + K::FE::TWSIZE + K::XE::TWSIZE);
    const long ntiles = (fp.nblocks + BF - 1) / BF;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_chan_pingpong<N, BF, EF, L, EX>, 2 * K::T, smem, ntiles, 1, &grid));
    return launch_tile_kernel(k_chan_pingpong<N, BF, EF, L, EX>, grid, 2 * K::T, smem, s, fp, xp, twiddle_table(N, EF), twiddle_table(L, EX), ntiles);
}

template <int N, int BF, int EF, int L, int EX> static cudaError_t go_fused(const FwdParams& fp, const ExtractParams& xp, cudaStream_t s)
{
    typedef ChanFused<N, BF, EF, L, EX> K;
    const long ntiles = (fp.nblocks + BF - 1) / BF;
    unsigned grid = 1;
    FDC_CHECK(persistent_grid(k_chan_fused<N, BF, EF, L, EX>, K::T, K::SMEM_BYTES, ntiles, 1, &grid));
    return launch_tile_kernel(k_chan_fused<N, BF, EF, L, EX>, grid, K::T, K::SMEM_BYTES, s, fp, xp, twiddle_table(N, EF), twiddle_table(L, EX), ntiles);
}

/* (N, l) pairs with a fused kernel: the BASELINE configurations with N <= 16384 and their neighbours */
#define FDC_FUSED_TABLE(X) \
    X(1024, 4, 16, 64, 16) X(1024, 4, 16, 128, 16) X(2048, 2, 16, 128, 16) X(2048, 2, 16, 256, 16) \
    X(4096, 1, 16, 128, 16) X(4096, 1, 16, 256, 16) X(4096, 1, 16, 512, 16) \
    X(8192, 1, 32, 128, 16) X(8192, 1, 32, 256, 16) X(8192, 1, 32, 512, 16) X(8192, 1, 32, 1024, 16) \
    X(16384, 1, 32, 256, 16) X(16384, 1, 32, 512, 16) X(16384, 1, 32, 1024, 16)

bool chan_fused_supported(int N, int l)
{
#define X(NN, BF, EF, LL, EX) if (N == NN && l == LL) return true;
    FDC_FUSED_TABLE(X)
#undef X
    return false;
}
cudaError_t launch_chan_fused(const FwdParams& fp, const ExtractParams& xp, int l, cudaStream_t s)
{
    if (fp.nblocks <= 0) return cudaSuccess;
    if (tuning().fuse_small > 1) {
#define X(NN, BF, EF, LL, EX) if constexpr (2 * ChanFused<NN, BF, EF, LL, EX>::T <= 1024 && sizeof(float2) * (2 * ChanFused<NN, BF, EF, LL, EX>::R1_ELEMS + ChanFused<NN, BF, EF, LL, EX>::R2_ELEMS + 2048) <= 227 * 1024) { if (fp.N == NN && l == LL) return go_pingpong<NN, BF, EF, LL, EX>(fp, xp, s); }
        FDC_FUSED_TABLE(X)
#undef X
    }
#define X(NN, BF, EF, LL, EX) if (fp.N == NN && l == LL) return go_fused<NN, BF, EF, LL, EX>(fp, xp, s);
    FDC_FUSED_TABLE(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace fdc
