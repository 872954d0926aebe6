/* fdc_functors.cuh -- the global-memory sides of the tile FFT: what each kernel fuses into its
 * first-pass loads and last-pass stores.  Host/device code (see fdc_hd.h).
 *
 * Every functor has   Ctx begin(batch)   -- per-signal addressing (block / channel / job lookup, phase selection,
 *                                            row pointers), evaluated once per butterfly, and
 *                     get(ctx, n) / put(ctx, k, v)  -- the per-element access, index arithmetic only.
 *
 * HBM layout
 *   input stream  : contiguous cfp32 samples of this call, preceded logically by `hist`
 *                   (the last ovl samples of the previous call, zeros at stream start --
 *                   lib/overlap_save_impl.cc:52,70-78)
 *   spectrum      : [block][N] cfp32, fft-shifted (DC at N/2) and scaled by 1/N -- exactly what
 *                   the hier block's normalize_input emits (python/FrequencyDomainChannelizer.py:206,216)
 *   tables        : per channel [phase][l] cfp32 (lib/windows.h:41-78)
 *   outputs       : channel-major slabs, channel c at out + nblocks_call * lout_prefix[c]      */
#ifndef FDC_FUNCTORS_CUH
#define FDC_FUNCTORS_CUH
#include "fdc_hd.h"

namespace fdc {

/* overlap-save window of one block: element n is in[blk*hop - ovl + n]; the first `nhist` elements of the very
 * first blocks of a launch come from the saved history instead (lib/overlap_save_impl.cc:70-78) */
struct OvlCtx { const float2* base; const float2* hbase; int nhist; };
FDC_HD OvlCtx ovl_ctx(const float2* in, const float2* hist, long blk, int hop, int ovl, bool valid)
{
    OvlCtx c;
    const long start = blk * hop - ovl;                 /* may be negative for the first blocks */
    c.base = in + start; c.hbase = hist + blk * hop;
    c.nhist = start < 0 ? (int)(-start) : 0;
    if (!valid) { c.base = 0; c.nhist = -1; }
    return c;
}
FDC_HD float2 ovl_get(const OvlCtx& c, long n)
{
    if (c.nhist < 0) return make_float2(0.f, 0.f);
    return n < c.nhist ? fdc_ldg(c.hbase + n) : fdc_ldg(c.base + n);
}

/* ------------------------------------------------------------------ K1: forward FFT, N in one CTA */
struct FwdParams {
    const float2* in;      /* new samples of this launch: block b starts at in[b*hop - ovl] */
    const float2* hist;    /* ovl samples preceding in[0] */
    float2* spec;          /* [nblocks][N] */
    long nblocks;
    int hop, ovl, N;
    float scale;           /* 1/N (a power of two: exact) */
};
template <int N, int B> struct FwdLoader {
    typedef OvlCtx Ctx;
    const FwdParams& p; int tile;
    FDC_HD Ctx begin(int batch) const
    {
        const long blk = (long)tile * B + batch;
        return ovl_ctx(p.in, p.hist, blk, p.hop, p.ovl, blk < p.nblocks);
    }
    FDC_HD float2 get(const Ctx& c, int n) const { return ovl_get(c, n); }
};
template <int N, int B> struct FwdStorer {
    typedef float2* Ctx;
    const FwdParams& p; int tile;
    FDC_HD Ctx begin(int batch) const
    {
        const long blk = (long)tile * B + batch;
        return blk < p.nblocks ? p.spec + blk * N : (float2*)0;
    }
    FDC_HD void put(const Ctx& row, int k, float2 v) const
    {
        /* fft_vcc shift=True for a forward transform: out[0:N/2] = Y[N/2:N], out[N/2:N] = Y[0:N/2] */
        if (row) row[k ^ (N / 2)] = make_float2(v.x * p.scale, v.y * p.scale);
    }
};

/* ------------------------------------------------------ K1 (large N = N1*N2): four-step, two kernels
 * n = N2*n1 + n2, k = k1 + N1*k2:
 *   pass A (columns): for every n2   A[k1][n2] = W_N^{n2 k1} * sum_n1 x[N2 n1 + n2] W_N1^{n1 k1}
 *   pass B (rows)   : for every k1   X[k1 + N1 k2] = sum_n2 A[k1][n2] W_N2^{n2 k2}
 * Both kernels work on tiles of 16 adjacent columns / rows so that every global access is a full
 * 128-byte line.  W_N^m is formed from two short tables: W_N^m = twlo[m & (TWS-1)] * twhi[m >> log2 TWS]. */
struct BigParams {
    const float2* in; const float2* hist;
    float2* mid;           /* [nblocks][N1][N2] intermediate */
    float2* spec;          /* [nblocks][N] */
    const float2* twlo; const float2* twhi; int tws_log2;
    long nblocks;
    int hop, ovl;
    float scale;
};
template <int N1, int N2, int B> struct ColLoader {     /* signal = column n2, element index = n1 */
    struct Ctx { OvlCtx o; int n2; };
    const BigParams& p; int tile; long blk;
    FDC_HD Ctx begin(int batch) const
    {
        Ctx c; c.n2 = tile * B + batch; c.o = ovl_ctx(p.in, p.hist, blk, p.hop, p.ovl, true);
        return c;
    }
    FDC_HD float2 get(const Ctx& c, int n1) const { return ovl_get(c.o, (long)N2 * n1 + c.n2); }
};
template <int N1, int N2, int B> struct ColStorer {
    struct Ctx { float2* col; unsigned n2; };
    const BigParams& p; int tile; long blk;
    FDC_HD Ctx begin(int batch) const
    {
        Ctx c; c.n2 = (unsigned)(tile * B + batch); c.col = p.mid + blk * ((long)N1 * N2) + c.n2;
        return c;
    }
    FDC_HD void put(const Ctx& c, int k1, float2 v) const
    {
        const unsigned m = c.n2 * (unsigned)k1;                        /* < N1*N2 */
        const float2 w = cmul(fdc_ldg(p.twlo + (m & ((1u << p.tws_log2) - 1u))), fdc_ldg(p.twhi + (m >> p.tws_log2)));
        c.col[(long)k1 * N2] = cmul(v, w);
    }
};
template <int N1, int N2, int B> struct RowLoader {     /* signal = row k1, element index = n2 */
    typedef const float2* Ctx;
    const BigParams& p; int tile; long blk;
    FDC_HD Ctx begin(int batch) const { return p.mid + blk * ((long)N1 * N2) + (long)(tile * B + batch) * N2; }
    FDC_HD float2 get(const Ctx& row, int n2) const { return row[n2]; }
};
template <int N1, int N2, int B> struct RowStorer {
    struct Ctx { float2* spec; int k1; };
    const BigParams& p; int tile; long blk;
    FDC_HD Ctx begin(int batch) const
    {
        Ctx c; c.k1 = tile * B + batch; c.spec = p.spec + blk * ((long)N1 * N2);
        return c;
    }
    FDC_HD void put(const Ctx& c, int k2, float2 v) const
    {
        const int k = c.k1 + N1 * k2;
        c.spec[k ^ (N1 * N2 / 2)] = make_float2(v.x * p.scale, v.y * p.scale);
    }
};

/* ------------------------------------------------------------------ K2: batched channel extract
 * One work item = (channel c, block b):
 *   y = IFFT_l( halfswap( X_b[f_c : f_c+l] .* table_c[phase] ) )[l-lout :] * gain
 * vector_cut_vxx (lib/vector_cut_vxx_impl.cc:67-68) is the address arithmetic,
 * phase_shifting_windowing_vcc::work (lib/phase_shifting_windowing_vcc_impl.cc:81-82) the table multiply with
 * phase = (blocks seen so far * shift) mod nphase, fft_vcc(l, inverse, shift=True) the half swap + backward FFT,
 * the second vector_cut drops the first l-lout samples, multiply_const the gain.
 * A CTA handles B consecutive blocks of ONE channel so the table stays in L1 and the stores are one run. */
struct ChanDev {
    int f, lout, shift, pad0;
    long tab_off;          /* float2 offset of table_c[0][0] in `tables` */
    long lout_prefix;      /* sum of lout over the channels before this one */
    float gain; int pad1;
};
struct ExtractParams {
    const float2* spec; long spec_stride;   /* rows of the spectrum (ring) holding this chunk */
    const float2* tables;
    const ChanDev* chans;
    const int* sel;        /* channel indices handled by this launch (all share l) */
    float2* out;
    long nb;               /* blocks in this chunk */
    long call_blocks;      /* blocks of the whole call (slab size) */
    long call_blk0;        /* index of the chunk's first block inside the call */
    int glob_phase0;       /* (global index of the chunk's first block) mod nphase */
    int nphase;
};
template <int L, int B> struct ExtractLoader {
    struct Ctx { const float2* x; const float2* w; };
    const ExtractParams& p; int tile; int ysel;
    FDC_HD Ctx begin(int batch) const
    {
        Ctx c; c.x = 0; c.w = 0;
        const long b = (long)tile * B + batch;
        if (b >= p.nb) return c;
        const ChanDev& ch = p.chans[fdc_ldg(p.sel + ysel)];
        const unsigned np = (unsigned)p.nphase;
        const unsigned phase = ((((unsigned)p.glob_phase0 + ((unsigned)b % np)) % np) * (unsigned)ch.shift) % np;
        c.x = p.spec + b * p.spec_stride + ch.f;
        c.w = p.tables + ch.tab_off + (long)phase * L;
        return c;
    }
    FDC_HD float2 get(const Ctx& c, int n) const
    {
        if (!c.x) return make_float2(0.f, 0.f);
        const int m = (n + L / 2) & (L - 1);              /* fft_vcc inverse+shift: dst[n] = in[(n + l/2) mod l] */
        return cmul_exact(fdc_ldg(c.x + m), fdc_ldg(c.w + m));
    }
};
template <int L, int B> struct ExtractStorer {
    struct Ctx { float2* dst; int skip; float gain; };
    const ExtractParams& p; int tile; int ysel;
    FDC_HD Ctx begin(int batch) const
    {
        Ctx c; c.dst = 0; c.skip = 0; c.gain = 0.f;
        const long b = (long)tile * B + batch;
        if (b >= p.nb) return c;
        const ChanDev& ch = p.chans[fdc_ldg(p.sel + ysel)];
        c.skip = L - ch.lout; c.gain = ch.gain;
        c.dst = p.out + p.call_blocks * ch.lout_prefix + (p.call_blk0 + b) * ch.lout - c.skip;
        return c;
    }
    FDC_HD void put(const Ctx& c, int k, float2 v) const
    {
        if (c.dst && k >= c.skip) c.dst[k] = make_float2(v.x * c.gain, v.y * c.gain);
    }
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
    const JobParams& p; int tile;
    FDC_HD Ctx begin(int batch) const
    {
        Ctx c; c.x = 0; c.w = 0;
        const int ji = tile * B + batch;
        if (ji >= p.njobs) return c;
        const ExtractJob& jb = p.jobs[ji];
        c.x = (jb.row < 0 ? p.hist : p.spec + (long)jb.row * p.spec_stride) + jb.start;
        c.w = p.tables + jb.tab_off;
        return c;
    }
    FDC_HD float2 get(const Ctx& c, int n) const
    {
        if (!c.x) return make_float2(0.f, 0.f);
        const int m = (n + L / 2) & (L - 1);
        return cmul_exact(fdc_ldg(c.x + m), fdc_ldg(c.w + m));
    }
};
template <int L, int B> struct JobStorer {
    struct Ctx { float2* dst; int skip; };
    const JobParams& p; int tile;
    FDC_HD Ctx begin(int batch) const
    {
        Ctx c; c.dst = 0; c.skip = 0;
        const int ji = tile * B + batch;
        if (ji >= p.njobs) return c;
        const ExtractJob& jb = p.jobs[ji];
        c.skip = jb.skip; c.dst = p.out + jb.dst_off - jb.skip;
        return c;
    }
    FDC_HD void put(const Ctx& c, int k, float2 v) const { if (c.dst && k >= c.skip) c.dst[k] = v; }
};

/* ------------------------------------------------- plain batched FFT (fft_vcc stage replacement) */
struct PlainParams { const float2* in; float2* out; long nvec; int shift; };
template <int L, int B, int DIR> struct PlainLoader {
    typedef const float2* Ctx;
    const PlainParams& p; int tile;
    FDC_HD Ctx begin(int batch) const
    {
        const long v = (long)tile * B + batch;
        return v < p.nvec ? p.in + v * L : (const float2*)0;
    }
    FDC_HD float2 get(const Ctx& row, int n) const
    {
        if (!row) return make_float2(0.f, 0.f);
        const int m = (DIR < 0 && p.shift) ? ((n + L / 2) & (L - 1)) : n;
        return fdc_ldg(row + m);
    }
};
template <int L, int B, int DIR> struct PlainStorer {
    typedef float2* Ctx;
    const PlainParams& p; int tile;
    FDC_HD Ctx begin(int batch) const
    {
        const long v = (long)tile * B + batch;
        return v < p.nvec ? p.out + v * L : (float2*)0;
    }
    FDC_HD void put(const Ctx& row, int k, float2 v) const
    {
        if (!row) return;
        const int m = (DIR > 0 && p.shift) ? (k ^ (L / 2)) : k;
        row[m] = v;
    }
};

}  // namespace fdc
#endif
