/* fdc_functors.cuh -- the global-memory sides of the tile FFT: what each kernel fuses into its
 * first-pass loads and last-pass stores.  Host/device code (see fdc_hd.h).
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
    const FwdParams& p; int tile;
    FDC_HD float2 operator()(int batch, int n) const
    {
        const long blk = (long)tile * B + batch;
        if (blk >= p.nblocks) return make_float2(0.f, 0.f);
        const long g = blk * p.hop + n - p.ovl;
        return g < 0 ? fdc_ldg(p.hist + (p.ovl + g)) : fdc_ldg(p.in + g);
    }
};
template <int N, int B> struct FwdStorer {
    const FwdParams& p; int tile;
    FDC_HD void operator()(int batch, int k, float2 v) const
    {
        const long blk = (long)tile * B + batch;
        if (blk >= p.nblocks) return;
        /* fft_vcc shift=True for a forward transform: out[0:N/2] = Y[N/2:N], out[N/2:N] = Y[0:N/2] */
        p.spec[blk * N + (k ^ (N / 2))] = make_float2(v.x * p.scale, v.y * p.scale);
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
    const BigParams& p; int tile; long blk;
    FDC_HD float2 operator()(int batch, int n1) const
    {
        const int n2 = tile * B + batch;
        const long g = blk * p.hop + (long)N2 * n1 + n2 - p.ovl;
        return g < 0 ? fdc_ldg(p.hist + (p.ovl + g)) : fdc_ldg(p.in + g);
    }
};
template <int N1, int N2, int B> struct ColStorer {
    const BigParams& p; int tile; long blk;
    FDC_HD void operator()(int batch, int k1, float2 v) const
    {
        const int n2 = tile * B + batch;
        const unsigned m = (unsigned)n2 * (unsigned)k1;                 /* < N1*N2 */
        const float2 w = cmul(fdc_ldg(p.twlo + (m & ((1u << p.tws_log2) - 1u))), fdc_ldg(p.twhi + (m >> p.tws_log2)));
        p.mid[blk * ((long)N1 * N2) + (long)k1 * N2 + n2] = cmul(v, w);
    }
};
template <int N1, int N2, int B> struct RowLoader {     /* signal = row k1, element index = n2 */
    const BigParams& p; int tile; long blk;
    FDC_HD float2 operator()(int batch, int n2) const
    {
        const int k1 = tile * B + batch;
        return p.mid[blk * ((long)N1 * N2) + (long)k1 * N2 + n2];
    }
};
template <int N1, int N2, int B> struct RowStorer {
    const BigParams& p; int tile; long blk;
    FDC_HD void operator()(int batch, int k2, float2 v) const
    {
        const int k1 = tile * B + batch;
        const int k = k1 + N1 * k2;
        p.spec[blk * ((long)N1 * N2) + (k ^ (N1 * N2 / 2))] = make_float2(v.x * p.scale, v.y * p.scale);
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
    long glob_blk0;        /* global index of the chunk's first block (phase origin) */
    int nphase;
};
template <int L, int B> struct ExtractLoader {
    const ExtractParams& p; int tile; int ysel;
    FDC_HD float2 operator()(int batch, int n) const
    {
        const long b = (long)tile * B + batch;
        if (b >= p.nb) return make_float2(0.f, 0.f);
        const ChanDev& c = p.chans[fdc_ldg(p.sel + ysel)];
        const int phase = (int)((((p.glob_blk0 + b) % p.nphase) * c.shift) % p.nphase);
        const int m = (n + L / 2) & (L - 1);              /* fft_vcc inverse+shift: dst[n] = in[(n + l/2) mod l] */
        const float2 x = fdc_ldg(p.spec + b * p.spec_stride + c.f + m);
        const float2 w = fdc_ldg(p.tables + c.tab_off + (long)phase * L + m);
        return cmul_exact(x, w);
    }
};
template <int L, int B> struct ExtractStorer {
    const ExtractParams& p; int tile; int ysel;
    FDC_HD void operator()(int batch, int k, float2 v) const
    {
        const long b = (long)tile * B + batch;
        if (b >= p.nb) return;
        const ChanDev& c = p.chans[fdc_ldg(p.sel + ysel)];
        const int skip = L - c.lout;
        if (k < skip) return;
        p.out[p.call_blocks * c.lout_prefix + (p.call_blk0 + b) * c.lout + (k - skip)] = make_float2(v.x * c.gain, v.y * c.gain);
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
    const JobParams& p; int tile;
    FDC_HD float2 operator()(int batch, int n) const
    {
        const int ji = tile * B + batch;
        if (ji >= p.njobs) return make_float2(0.f, 0.f);
        const ExtractJob& jb = p.jobs[ji];
        const int m = (n + L / 2) & (L - 1);
        const float2* row = jb.row < 0 ? p.hist : p.spec + (long)jb.row * p.spec_stride;
        return cmul_exact(fdc_ldg(row + jb.start + m), fdc_ldg(p.tables + jb.tab_off + m));
    }
};
template <int L, int B> struct JobStorer {
    const JobParams& p; int tile;
    FDC_HD void operator()(int batch, int k, float2 v) const
    {
        const int ji = tile * B + batch;
        if (ji >= p.njobs) return;
        const ExtractJob& jb = p.jobs[ji];
        if (k < jb.skip) return;
        p.out[jb.dst_off + (k - jb.skip)] = v;
    }
};

/* ------------------------------------------------- plain batched FFT (fft_vcc stage replacement) */
struct PlainParams { const float2* in; float2* out; long nvec; int shift; };
template <int L, int B, int DIR> struct PlainLoader {
    const PlainParams& p; int tile;
    FDC_HD float2 operator()(int batch, int n) const
    {
        const long v = (long)tile * B + batch;
        if (v >= p.nvec) return make_float2(0.f, 0.f);
        const int m = (DIR < 0 && p.shift) ? ((n + L / 2) & (L - 1)) : n;
        return fdc_ldg(p.in + v * L + m);
    }
};
template <int L, int B, int DIR> struct PlainStorer {
    const PlainParams& p; int tile;
    FDC_HD void operator()(int batch, int k, float2 v) const
    {
        const long vv = (long)tile * B + batch;
        if (vv >= p.nvec) return;
        const int m = (DIR > 0 && p.shift) ? (k ^ (L / 2)) : k;
        p.out[vv * L + m] = v;
    }
};

}  // namespace fdc
#endif
