/* fdc_functors.cuh -- the global-memory sides of the tile FFT: what each kernel fuses into its
 * first-pass loads and last-pass stores.  Host/device code (see fdc_hd.h).
 *
 * Loader :  Ctx begin(batch, j)                    -- one address computation per butterfly
 *           fetch<R, STRIDE>(ctx, t)               -- raw load of FFT input element j + t*STRIDE (t unrolled, so the
 *                                                     access is base + immediate); issued one tile ahead (prefetch)
 *           finish<R, STRIDE>(ctx, t, raw)         -- arithmetic on the fetched value (HAS_FINISH)
 * Storer :  Ctx begin(batch, o)                    -- o < NS
 *           put<R, NS>(ctx, t, v)                  -- FFT output element o + t*NS
 * Half swaps (fft_vcc shift=True) are static permutations of t: L/2 = (R/2) * STRIDE on the load side and
 * (R/2) * NS on the store side.
 *
 * HBM layout
 *   input stream  : contiguous cfp32 samples; block b reads in[b*hop - ovl, b*hop + hop): the ovl samples
 *                   before in[0] must be addressable (the C-ABI layer stages [history | first blocks] for the
 *                   blocks that reach back into the previous call -- lib/overlap_save_impl.cc:52,70-78)
 *   spectrum      : [block][N] cfp32, fft-shifted (DC at N/2) and scaled by 1/N -- exactly what
 *                   the hier block's normalize_input emits (python/FrequencyDomainChannelizer.py:206,216)
 *   tables        : [phase][l] cfp32 per distinct table (lib/windows.h:41-78); channels with equal tables share one
 *   outputs       : channel-major slabs, channel c at out + nblocks_call * lout_prefix[c]      */
#ifndef FDC_FUNCTORS_CUH
#define FDC_FUNCTORS_CUH
#include "fdc_hd.h"
#include "fdc_tma.cuh"

namespace fdc {

/* tile index = outer * ninner + inner (inner runs fastest).  Persistent loops split their first tile and their stride
 * once and then advance the pair without dividing. */
struct TilePos { int inner; int outer; };
FDC_HD TilePos tile_split(long tile, int ninner)
{
    TilePos s; s.outer = (int)(tile / ninner); s.inner = (int)(tile - (long)s.outer * ninner);
    return s;
}
FDC_HD TilePos tile_advance(TilePos a, TilePos step, int ninner)
{
    a.inner += step.inner; a.outer += step.outer;
    if (a.inner >= ninner) { a.inner -= ninner; a.outer++; }
    return a;
}

/* ------------------------------------------------------------------ K1: forward FFT, N in one CTA */
struct FwdParams {
    const float2* in;      /* block b is in[b*hop - ovl, b*hop + hop) */
    float2* spec;          /* [nblocks][N] */
    long nblocks;
    int hop, ovl, N;
    float scale;           /* 1/N (a power of two: exact) */
    int l2pf;              /* kernels without a register prefetch: bulk-prefetch the next tile's samples into L2 */
    /* overlap-save history (lib/overlap_save_impl.cc:70-78): the first head_blocks blocks reach back before in[0]; those
     * samples are hist[ovl + i] for sample index i < 0 (hist = the last ovl samples of the previous call, zeros at the start).
     * Tiles that hold such a block take the two-segment loads, all others the plain ones (a tile-uniform branch).
     * head_off: samples of the same call that lie in front of in[0] (a later chunk of the call: in[-head_off .. 0) is the
     * caller's buffer, the history starts before that). */
    const float2* hist; long head_blocks; long head_off;
};
template <int N, int B> struct FwdLoader {
    typedef const float2* Ctx;
    static constexpr bool HAS_FINISH = false;
    const FwdParams& p; long tile;
    FDC_HD Ctx begin(int batch, int j) const
    {
        /* a ragged last tile re-reads the last block (its results are not stored) */
        long blk = tile * B + batch;
        if (blk >= p.nblocks) blk = p.nblocks - 1;
        return p.in + (blk * p.hop - p.ovl + j);
    }
    template <int R, int STRIDE> FDC_HD float2 fetch(const Ctx& c, int t) const { return fdc_ldg(c + t * STRIDE); }
    template <int R, int STRIDE> FDC_HD float2 finish(const Ctx&, int, float2 raw) const { return raw; }
    static constexpr bool HAS_HEAD = true;
    FDC_HD bool is_head() const { return tile * B < p.head_blocks; }
    template <int R, int STRIDE> FDC_HD float2 fetch_head(const Ctx& c, int t) const
    {
        const long i = (c - p.in) + t * STRIDE;            /* sample index relative to in[0] */
        return i < -p.head_off ? p.hist[p.ovl + i + p.head_off] : fdc_ldg(p.in + i);
    }
};
template <int N, int B> struct FwdStorer {
    typedef float2* Ctx;
    const FwdParams& p; long tile;
    FDC_HD Ctx begin(int batch, int o) const
    {
        const long blk = tile * B + batch;
        return blk < p.nblocks ? p.spec + (blk * N + o) : (float2*)0;
    }
    template <int R, int NS> FDC_HD void put(const Ctx& row, int t, float2 v) const
    {
        /* fft_vcc shift=True for a forward transform: out[0:N/2] = Y[N/2:N], out[N/2:N] = Y[0:N/2] */
        if (row) row[(t ^ (R / 2)) * NS] = cscale(v, p.scale);
    }
};
template <int N, int B> struct FwdTiles {
    const FwdParams& p;
    FDC_HD int ninner() const { return 1; }
    FDC_HD FwdLoader<N, B> loader(TilePos t) const { return FwdLoader<N, B>{p, (long)t.outer}; }
    FDC_HD FwdStorer<N, B> storer(TilePos t) const { return FwdStorer<N, B>{p, (long)t.outer}; }
#if defined(__CUDACC__) && !defined(FDC_HOST_EMU)
    /* the B blocks of a tile are one contiguous run of samples: one bulk prefetch by one thread */
    static constexpr bool HAS_L2_PREFETCH = true;
    __device__ __forceinline__ void prefetch_l2(TilePos t, int tid) const
    {
        if (!p.l2pf || tid != 0) return;
        const long blk0 = (long)t.outer * B;
        if (blk0 < p.head_blocks) return;                  /* part of these samples lives in the history buffer */
        const long nb = p.nblocks - blk0 < B ? p.nblocks - blk0 : B;
        if (nb <= 0) return;
        const char* a = reinterpret_cast<const char*>(p.in + (blk0 * p.hop - p.ovl));
        long bytes = (long)sizeof(float2) * ((nb - 1) * p.hop + N);
        if (reinterpret_cast<uintptr_t>(a) & 15) { a += 8; bytes -= 8; }          /* float2 granularity: at most 8 bytes off */
        bytes &= ~15L;
        if (bytes > 0) bulk_prefetch_l2(a, (uint32_t)bytes);
    }
#endif
};

/* ------------------------------------------------------ K1 (large N = N1*N2): four-step, two kernels
 * n = N2*n1 + n2, k = k1 + N1*k2:
 *   pass A (columns): for every n2   A[k1][n2] = W_N^{n2 k1} * sum_n1 x[N2 n1 + n2] W_N1^{n1 k1}
 *   pass B (rows)   : for every k1   X[k1 + N1 k2] = sum_n2 A[k1][n2] W_N2^{n2 k2}
 * Both kernels work on tiles of 16 adjacent columns / rows so that every global access is a full
 * 128-byte line.  W_N^{n2 k1} comes from a table laid out like the intermediate ([k1][n2], N entries,
 * L2 resident); a persistent CTA keeps one column tile, so its 32 KB of the table stay in L1. */
struct BigParams {
    const float2* in;
    float2* mid;           /* [nblocks][N1][N2] intermediate */
    float2* spec;          /* [nblocks][N] */
    const float2* tw4;     /* [N1][N2] four-step twiddles */
    long nblocks;
    int hop, ovl;
    float scale;
    const float2* hist; long head_blocks; long head_off;     /* as in FwdParams */
};
template <int N1, int N2, int B> struct ColLoader {     /* signal = column n2, element index = n1 */
    typedef const float2* Ctx;
    static constexpr bool HAS_FINISH = false;
    const BigParams& p; int ctile; long blk;
    FDC_HD Ctx begin(int batch, int j) const { return p.in + (blk * p.hop - p.ovl + (long)N2 * j + (ctile * B + batch)); }
    template <int R, int STRIDE> FDC_HD float2 fetch(const Ctx& c, int t) const { return fdc_ldg(c + (long)t * STRIDE * N2); }
    template <int R, int STRIDE> FDC_HD float2 finish(const Ctx&, int, float2 raw) const { return raw; }
    static constexpr bool HAS_HEAD = true;
    FDC_HD bool is_head() const { return blk < p.head_blocks; }
    template <int R, int STRIDE> FDC_HD float2 fetch_head(const Ctx& c, int t) const
    {
        const long i = (c - p.in) + (long)t * STRIDE * N2;
        return i < -p.head_off ? p.hist[p.ovl + i + p.head_off] : fdc_ldg(p.in + i);
    }
};
/* The four-step twiddles a column CTA needs are the same for every block (it keeps its column tile), so they are
 * copied once into shared memory, laid out [t][butterfly] = exactly the order the last pass consumes them: every
 * thread reads back only the 16 entries it wrote itself (conflict free, no barrier, nothing left in L1/L2 traffic). */
template <int N1, int N2, int B> struct ColTwiddles {
    template <int R, int NS> static FDC_HD void init_one(float2* tws, const float2* tw4, int ctile, int batch, int o)
    {
        const float2* src = tw4 + ((long)o * N2 + (ctile * B + batch));
#pragma unroll
        for (int t = 0; t < R; t++) tws[t * (NS * B) + o * B + batch] = fdc_ldg(src + (long)t * NS * N2);   /* NS * B butterflies per tile */
    }
};
template <int N1, int N2, int B> struct ColStorer {
    struct Ctx { float2* col; const float2* tw; };
    const BigParams& p; int ctile; long blk; const float2* tws;
    FDC_HD Ctx begin(int batch, int o) const
    {
        Ctx c; c.col = p.mid + (blk * ((long)N1 * N2) + (long)o * N2 + (ctile * B + batch)); c.tw = tws + (o * B + batch);
        return c;
    }
    template <int R, int NS> FDC_HD void put(const Ctx& c, int t, float2 v) const
    {
        c.col[(long)t * NS * N2] = cmul(v, c.tw[t * (NS * B)]);
    }
};
template <int N1, int N2, int B> struct ColTiles {      /* tile = blk * (N2/B) + column tile */
    const BigParams& p; const float2* tws;
    FDC_HD int ninner() const { return N2 / B; }
    FDC_HD ColLoader<N1, N2, B> loader(TilePos t) const { return ColLoader<N1, N2, B>{p, t.inner, (long)t.outer}; }
    FDC_HD ColStorer<N1, N2, B> storer(TilePos t) const { return ColStorer<N1, N2, B>{p, t.inner, (long)t.outer, tws}; }
};
template <int N1, int N2, int B> struct RowLoader {     /* signal = row k1, element index = n2 */
    typedef const float2* Ctx;
    static constexpr bool HAS_FINISH = false;
    const BigParams& p; int rtile; long blk;
    FDC_HD Ctx begin(int batch, int j) const { return p.mid + (blk * ((long)N1 * N2) + (long)(rtile * B + batch) * N2 + j); }
    template <int R, int STRIDE> FDC_HD float2 fetch(const Ctx& row, int t) const { return row[t * STRIDE]; }
    template <int R, int STRIDE> FDC_HD float2 finish(const Ctx&, int, float2 raw) const { return raw; }
};
template <int N1, int N2, int B> struct RowStorer {
    typedef float2* Ctx;
    const BigParams& p; int rtile; long blk;
    FDC_HD Ctx begin(int batch, int o) const { return p.spec + (blk * ((long)N1 * N2) + (rtile * B + batch) + (long)N1 * o); }
    template <int R, int NS> FDC_HD void put(const Ctx& c, int t, float2 v) const
    {
        /* k = k1 + N1 (o + t NS); k ^ (N/2) flips bit log2(N2/2) of k2, i.e. t -> t ^ (R/2) */
        c[(long)N1 * NS * (t ^ (R / 2))] = cscale(v, p.scale);
    }
};
template <int N1, int N2, int B> struct RowTiles {      /* tile = blk * (N1/B) + row tile */
    const BigParams& p;
    FDC_HD int ninner() const { return N1 / B; }
    FDC_HD RowLoader<N1, N2, B> loader(TilePos t) const { return RowLoader<N1, N2, B>{p, t.inner, (long)t.outer}; }
    FDC_HD RowStorer<N1, N2, B> storer(TilePos t) const { return RowStorer<N1, N2, B>{p, t.inner, (long)t.outer}; }
};

/* ------------------------------------------------------------------ K2: batched channel extract
 * One work item = (channel c, block b):
 *   y = IFFT_l( halfswap( X_b[f_c : f_c+l] .* table_c[phase] ) )[l-lout :] * gain
 * vector_cut_vxx (lib/vector_cut_vxx_impl.cc:67-68) is the address arithmetic,
 * phase_shifting_windowing_vcc::work (lib/phase_shifting_windowing_vcc_impl.cc:81-82) the table multiply with
 * phase = (blocks seen so far * shift) mod nphase, fft_vcc(l, inverse, shift=True) the half swap + backward FFT,
 * the second vector_cut drops the first l-lout samples, multiply_const the gain.
 * A CTA handles B channels that are neighbours in frequency on ONE block: their slices overlap, so the spectrum is
 * read from L2 once and the second use hits in L1. */
struct ChanDev {
    int f, lout, shift, owner;   /* owner: index of the sink this channel's rows go to (channel-sharded sinks), else 0 */
    long tab_off;          /* float2 offset of table[0][0] in `tables` */
    long lout_prefix;      /* sum of lout over the channels before this one */
    long sink_prefix;      /* sum of lout over the earlier channels of the same owner */
    float gain; int pad1;
};
#define FDC_MAX_SINKS 16
struct ExtractParams {
    const float2* spec; long spec_stride;   /* rows of the spectrum (ring) holding this chunk */
    const float2* tables;
    const ChanDev* chans;  /* the nsel channels handled by this launch (all share l), ascending f */
    int nsel, ny;          /* ny = ceil(nsel / B) channel tiles */
    float2* out;
    long nb;               /* blocks in this chunk */
    long call_blocks;      /* blocks of the whole call (slab size) */
    long call_blk0;        /* index of the chunk's first block inside the call */
    int glob_phase0;       /* (global index of the chunk's first block) mod nphase */
    int nphase;
    int tma_ok;            /* every slice of this launch is 16-byte aligned (even f, aligned spectrum base and stride) */
    int l2pf;              /* bulk-prefetch the next tile's slices into L2 while this tile is transformed (needs tma_ok) */
    int bpt;               /* packed tiles (few channels of this length): a tile holds all nsel channels of bpt consecutive blocks */
    int phase_mask;        /* nphase - 1 when nphase is a power of two (the hier block's relinvovl always is), else -1 */
    /* channel-sharded sinks (multi-GPU, FDC/sharded.py ChannelSinks): channel c's rows go to sink[owner_c] -- the buffer of the
     * GPU that owns the channel (own memory, or peer memory mapped over NVLink), or a local staging slab that the copy engines
     * forward to the owner.  Every sink is laid out like `out` but holds only its owner's channels: blocks rows per channel,
     * this launch's blocks starting at row blk0.  nsinks == 0: everything goes to `out`. */
    int nsinks;
    struct Sink { float2* base; long blocks, blk0; } sink[FDC_MAX_SINKS];
};
/* first item of block b of channel ch in its output buffer */
FDC_HD float2* chan_out_row(const ExtractParams& p, const ChanDev& ch, long b)
{
    if (p.nsinks) {
        const ExtractParams::Sink& k = p.sink[ch.owner];
        return k.base + (k.blocks * ch.sink_prefix + (k.blk0 + b) * ch.lout);
    }
    return p.out + (p.call_blocks * ch.lout_prefix + (p.call_blk0 + b) * ch.lout);
}
/* x mod nphase without an integer division when nphase is a power of two */
FDC_HD unsigned phase_mod(const ExtractParams& p, unsigned x) { return p.phase_mask >= 0 ? (x & (unsigned)p.phase_mask) : x % (unsigned)p.nphase; }
template <int L, int B> struct ExtractLoader {
    struct Ctx { const float2* x; const float2* w; };
    static constexpr bool HAS_FINISH = true;
    const ExtractParams& p; int ytile; long b; unsigned bphase;     /* bphase = (global block index) mod nphase */
    FDC_HD Ctx begin(int batch, int j) const
    {
        Ctx c;
        int s = ytile * B + batch;
        if (s >= p.nsel) s = p.nsel - 1;                   /* ragged last tile: recompute the last channel, store nothing */
        const ChanDev& ch = p.chans[s];
        const unsigned phase = phase_mod(p, bphase * (unsigned)ch.shift);
        c.x = p.spec + (b * p.spec_stride + ch.f + j);
        c.w = p.tables + (ch.tab_off + (long)phase * L + j);
        return c;
    }
    /* fft_vcc inverse+shift: FFT input n takes bin (n + l/2) mod l; n = j + t STRIDE, l/2 = (R/2) STRIDE */
    template <int R, int STRIDE> FDC_HD float2 fetch(const Ctx& c, int t) const { return fdc_ldg(c.x + ((t + R / 2) % R) * STRIDE); }
    /* fused path: contracted complex multiply (2 FMUL + 2 FFMA); the stand-alone phase_shifting_windowing_vcc block and
     * the activity-gated job path keep the VOLK-exact, uncontracted form */
    template <int R, int STRIDE> FDC_HD float2 finish(const Ctx& c, int t, float2 raw) const
    {
        return cmul(raw, fdc_ldg(c.w + ((t + R / 2) % R) * STRIDE));
    }
};
template <int L, int B> struct ExtractStorer {
    struct Ctx { float2* dst; int skip; float gain; };
    const ExtractParams& p; int ytile; long b;
    FDC_HD Ctx begin(int batch, int o) const
    {
        Ctx c; c.dst = p.out; c.skip = 1 << 30; c.gain = 0.f;          /* skip beyond L: nothing is stored */
        const int s = ytile * B + batch;
        if (s >= p.nsel) return c;
        const ChanDev& ch = p.chans[s];
        c.skip = L - ch.lout - o; c.gain = ch.gain;
        c.dst = chan_out_row(p, ch, b) - (L - ch.lout) + o;
        return c;
    }
    /* gain 1 (a power-of-two gain is folded into the table at create time) skips the multiply: uniform branch per butterfly */
    static constexpr bool HAS_VARIANT = true;
    FDC_HD bool variant(const Ctx& c) const { return c.gain == 1.0f; }
    template <int R, int NS, bool UNIT> FDC_HD void put(const Ctx& c, int t, float2 v) const
    {
        if (t * NS >= c.skip) c.dst[t * NS] = UNIT ? v : cscale(v, c.gain);
    }
};
template <int L, int B> struct ExtractTiles {           /* tile = block * ny + channel tile */
    const ExtractParams& p;
    FDC_HD int ninner() const { return p.ny; }
    FDC_HD ExtractLoader<L, B> loader(TilePos t) const
    {
        return ExtractLoader<L, B>{p, t.inner, (long)t.outer, phase_mod(p, (unsigned)p.glob_phase0 + (unsigned)t.outer)};
    }
    FDC_HD ExtractStorer<L, B> storer(TilePos t) const { return ExtractStorer<L, B>{p, t.inner, (long)t.outer}; }
#if defined(__CUDACC__) && !defined(FDC_HOST_EMU)
    /* ONE thread asks the TMA engine to pull the tile's slices into L2: the channels of a tile are neighbours in frequency
     * (ascending f), so their slices are normally one contiguous run of bins and a single instruction covers them; a
     * sparse group falls back to one request per channel.  (Per-lane requests would be serialised lane by lane: the
     * instruction takes its operands from uniform registers.) */
    static constexpr bool HAS_L2_PREFETCH = true;
    __device__ __forceinline__ void prefetch_l2(TilePos t, int tid) const
    {
        if (tid != 0 || !p.l2pf) return;
        const int s0 = t.inner * B;
        const int s1 = (s0 + B <= p.nsel ? s0 + B : p.nsel) - 1;
        const float2* row = p.spec + (long)t.outer * p.spec_stride;
        const int f0 = p.chans[s0].f, span = p.chans[s1].f - f0 + L;
        if (span <= 2 * B * L) bulk_prefetch_l2(row + f0, (uint32_t)(sizeof(float2) * span));
        else {
#pragma unroll 1
            for (int s = s0; s <= s1; s++) bulk_prefetch_l2(row + p.chans[s].f, (uint32_t)(sizeof(float2) * L));
        }
    }
#endif
};

/* Packed tiles: when a launch has fewer than B / 2 channels of this length (the usual GNU Radio flowgraph has a handful),
 * one block does not fill a tile and most of the CTA would transform padding.  A packed tile takes all nsel channels of
 * bpt = B / nsel consecutive blocks instead: signal `batch` is channel batch % nsel of block first + batch / nsel. */
template <int L, int B> struct PackedSignal {
    long blk; int s; bool valid;
    FDC_HD PackedSignal(const ExtractParams& p, long first, int batch)
    {
        const int db = batch / p.nsel;
        s = batch - db * p.nsel; blk = first + db;
        valid = db < p.bpt && blk < p.nb;
        if (!valid) { blk = first; s = 0; }                /* padding signals recompute the first one, nothing is stored */
    }
};
template <int L, int B> struct PackedExtractLoader {
    struct Ctx { const float2* x; const float2* w; };
    static constexpr bool HAS_FINISH = true;
    const ExtractParams& p; long first;
    FDC_HD Ctx begin(int batch, int j) const
    {
        Ctx c;
        const PackedSignal<L, B> g(p, first, batch);
        const ChanDev& ch = p.chans[g.s];
        const unsigned phase = phase_mod(p, phase_mod(p, (unsigned)p.glob_phase0 + (unsigned)g.blk) * (unsigned)ch.shift);
        c.x = p.spec + (g.blk * p.spec_stride + ch.f + j);
        c.w = p.tables + (ch.tab_off + (long)phase * L + j);
        return c;
    }
    template <int R, int STRIDE> FDC_HD float2 fetch(const Ctx& c, int t) const { return fdc_ldg(c.x + ((t + R / 2) % R) * STRIDE); }
    template <int R, int STRIDE> FDC_HD float2 finish(const Ctx& c, int t, float2 raw) const
    {
        return cmul(raw, fdc_ldg(c.w + ((t + R / 2) % R) * STRIDE));
    }
};
template <int L, int B> struct PackedExtractStorer {
    struct Ctx { float2* dst; int skip; float gain; };
    const ExtractParams& p; long first;
    FDC_HD Ctx begin(int batch, int o) const
    {
        Ctx c; c.dst = p.out; c.skip = 1 << 30; c.gain = 0.f;
        const PackedSignal<L, B> g(p, first, batch);
        if (!g.valid) return c;
        const ChanDev& ch = p.chans[g.s];
        c.skip = L - ch.lout - o; c.gain = ch.gain;
        c.dst = chan_out_row(p, ch, g.blk) - (L - ch.lout) + o;
        return c;
    }
    static constexpr bool HAS_VARIANT = true;
    FDC_HD bool variant(const Ctx& c) const { return c.gain == 1.0f; }
    template <int R, int NS, bool UNIT> FDC_HD void put(const Ctx& c, int t, float2 v) const
    {
        if (t * NS >= c.skip) c.dst[t * NS] = UNIT ? v : cscale(v, c.gain);
    }
};
template <int L, int B> struct PackedExtractTiles {     /* tile = group of bpt blocks */
    const ExtractParams& p;
    FDC_HD int ninner() const { return 1; }
    FDC_HD PackedExtractLoader<L, B> loader(TilePos t) const { return PackedExtractLoader<L, B>{p, (long)t.outer * p.bpt}; }
    FDC_HD PackedExtractStorer<L, B> storer(TilePos t) const { return PackedExtractStorer<L, B>{p, (long)t.outer * p.bpt}; }
#if defined(__CUDACC__) && !defined(FDC_HOST_EMU)
    static constexpr bool HAS_L2_PREFETCH = true;
    __device__ __forceinline__ void prefetch_l2(TilePos t, int tid) const
    {
        if (tid != 0 || !p.l2pf) return;
        const long first = (long)t.outer * p.bpt;
        const int f0 = p.chans[0].f, span = p.chans[p.nsel - 1].f - f0 + L;     /* all channels of the launch, ascending f */
#pragma unroll 1
        for (int db = 0; db < p.bpt && first + db < p.nb; db++) {
            const float2* row = p.spec + (first + db) * p.spec_stride;
            if (span <= 4 * p.nsel * L) bulk_prefetch_l2(row + f0, (uint32_t)(sizeof(float2) * span));
            else {
#pragma unroll 1
                for (int c = 0; c < p.nsel; c++) bulk_prefetch_l2(row + p.chans[c].f, (uint32_t)(sizeof(float2) * L));
            }
        }
    }
#endif
};

/* ------------------------------------------------- K2 (activity gated): explicit job list
 * process_channel of PowerActivationChannel / SegmentDetection / activity_detection_channelizer_vcm
 * (lib/PowerActivationChannel_impl.cc:260-284, lib/SegmentDetection_impl.cc:399-429,
 *  lib/activity_detection_channelizer_vcm_impl.cc:373-397): window multiply, fftshift, backward FFT, drop ovlskip. */
struct ExtractJob {
    int row;               /* spectrum row of this call; -1 = the saved history block */
    int start;             /* extract_start */
    int tab_off;           /* float2 offset of the window (already phase selected) in `tables` */
    int skip;              /* ovlskip / output_ovl_offset */
    long dst_off;          /* float2 offset in `out` */
};
struct JobParams {
    const float2* spec; long spec_stride; const float2* hist;
    const float2* tables; const ExtractJob* jobs; float2* out; int njobs;
};
template <int L, int B> struct JobLoader {
    struct Ctx { const float2* x; const float2* w; };
    static constexpr bool HAS_FINISH = true;
    const JobParams& p; long tile;
    FDC_HD Ctx begin(int batch, int j) const
    {
        Ctx c;
        long ji = tile * B + batch;
        if (ji >= p.njobs) ji = p.njobs - 1;
        const ExtractJob& jb = p.jobs[ji];
        c.x = (jb.row < 0 ? p.hist : p.spec + (long)jb.row * p.spec_stride) + (jb.start + j);
        c.w = p.tables + (jb.tab_off + j);
        return c;
    }
    template <int R, int STRIDE> FDC_HD float2 fetch(const Ctx& c, int t) const { return fdc_ldg(c.x + ((t + R / 2) % R) * STRIDE); }
    template <int R, int STRIDE> FDC_HD float2 finish(const Ctx& c, int t, float2 raw) const
    {
        return cmul_exact(raw, fdc_ldg(c.w + ((t + R / 2) % R) * STRIDE));
    }
};
template <int L, int B> struct JobStorer {
    struct Ctx { float2* dst; int skip; };
    const JobParams& p; long tile;
    FDC_HD Ctx begin(int batch, int o) const
    {
        Ctx c; c.dst = p.out; c.skip = 1 << 30;
        const long ji = tile * B + batch;
        if (ji >= p.njobs) return c;
        const ExtractJob& jb = p.jobs[ji];
        c.skip = jb.skip - o; c.dst = p.out + (jb.dst_off - jb.skip + o);
        return c;
    }
    template <int R, int NS> FDC_HD void put(const Ctx& c, int t, float2 v) const { if (t * NS >= c.skip) c.dst[t * NS] = v; }
};
template <int L, int B> struct JobTiles {
    const JobParams& p;
    FDC_HD int ninner() const { return 1; }
    FDC_HD JobLoader<L, B> loader(TilePos t) const { return JobLoader<L, B>{p, (long)t.outer}; }
    FDC_HD JobStorer<L, B> storer(TilePos t) const { return JobStorer<L, B>{p, (long)t.outer}; }
};

/* ------------------------------------------------- plain batched FFT (fft_vcc stage replacement) */
struct PlainParams { const float2* in; float2* out; long nvec; int shift; };
template <int L, int B, int DIR> struct PlainLoader {
    typedef const float2* Ctx;
    static constexpr bool HAS_FINISH = false;
    const PlainParams& p; long tile;
    FDC_HD Ctx begin(int batch, int j) const
    {
        long v = tile * B + batch;
        if (v >= p.nvec) v = p.nvec - 1;
        return p.in + (v * L + j);
    }
    template <int R, int STRIDE> FDC_HD float2 fetch(const Ctx& row, int t) const
    {
        const int tt = (DIR < 0 && p.shift) ? (t + R / 2) % R : t;       /* inverse + shift: half swap of the input */
        return fdc_ldg(row + tt * STRIDE);
    }
    template <int R, int STRIDE> FDC_HD float2 finish(const Ctx&, int, float2 raw) const { return raw; }
};
template <int L, int B, int DIR> struct PlainStorer {
    typedef float2* Ctx;
    const PlainParams& p; long tile;
    FDC_HD Ctx begin(int batch, int o) const
    {
        const long v = tile * B + batch;
        return v < p.nvec ? p.out + (v * L + o) : (float2*)0;
    }
    template <int R, int NS> FDC_HD void put(const Ctx& row, int t, float2 v) const
    {
        if (!row) return;
        const int tt = (DIR > 0 && p.shift) ? (t ^ (R / 2)) : t;         /* forward + shift: half swap of the output */
        row[tt * NS] = v;
    }
};
template <int L, int B, int DIR> struct PlainTiles {
    const PlainParams& p;
    FDC_HD int ninner() const { return 1; }
    FDC_HD PlainLoader<L, B, DIR> loader(TilePos t) const { return PlainLoader<L, B, DIR>{p, (long)t.outer}; }
    FDC_HD PlainStorer<L, B, DIR> storer(TilePos t) const { return PlainStorer<L, B, DIR>{p, (long)t.outer}; }
};

}  // namespace fdc
#endif
