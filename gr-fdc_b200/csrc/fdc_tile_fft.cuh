/* fdc_tile_fft.cuh -- CTA-level batched complex-fp32 FFT ("tile FFT").
 *
 * One CTA transforms a tile of B independent length-L signals (B*L = 16*T elements, T threads,
 * 16 elements per thread held in registers).  Stockham autosort passes with register radix-16/8/4/2
 * butterflies; between passes the tile is exchanged through shared memory (in place: read all ->
 * barrier -> write all -> barrier).  The first pass reads its operands straight from global memory
 * through a Loader functor (overlap-save addressing, window multiply, half swap ... are fused
 * there), the last pass hands its results to a Storer functor (fft-shift, overlap discard, scaling).
 *
 * Pass p (radix R, Ns = product of the earlier radices), butterfly j in [0, L/R):
 *      k    = j mod Ns
 *      in   : x[j + t*L/R] * W_{Ns*R}^{k t},  t = 0..R-1
 *      out  : y[(j-k)*R + k + t*Ns]
 * The first pass is radix 16 whenever L >= 16, so Ns >= 16 afterwards and every later exchange is
 * bank-conflict free as is; the first exchange (Ns = 1, thread writes 16 consecutive points) is
 * stored with one pad slot per 16 points (pos + pos/16).  The per-signal stride LP is odd so that
 * the "batch-fast" thread mappings (lanes walk over signals; used when the global side is a strided
 * column tile) stay conflict free too.
 *
 * The code is organised in PHASES separated by CTA barriers so that the same functions run on the
 * device and under the host emulator (see fdc_hd.h). */
#ifndef FDC_TILE_FFT_CUH
#define FDC_TILE_FFT_CUH
#include "fdc_bfly.cuh"

namespace fdc {

constexpr int fft_npasses(int L)
{
    if (L <= 16) return 1;
    int n = 1, rem = L / 16;
    while (rem > 1) {
        if (rem >= 64 || rem == 16) rem /= 16;
        else if (rem == 32) rem /= 8;
        else rem = 1;
        n++;
    }
    return n;
}
constexpr int fft_radix(int L, int p)
{
    if (L <= 16) return L;
    if (p == 0) return 16;
    int rem = L / 16, q = 1, r = 1;
    while (true) {
        if (rem >= 64 || rem == 16) r = 16;
        else if (rem == 32) r = 8;
        else r = rem;
        if (q == p) return r;
        rem /= r; q++;
    }
}
constexpr int fft_ns(int L, int p)
{
    int ns = 1;
    for (int q = 0; q < p; q++) ns *= fft_radix(L, q);
    return ns;
}
constexpr int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

template <int L_, int B_, int DIR, bool LOAD_BF, bool STORE_BF>
struct TileFFT {
    static constexpr int L = L_, B = B_, E = 16;
    static constexpr int T = L * B / E;                       /* threads per CTA */
    static constexpr int NP = fft_npasses(L);
    static constexpr int NPH = NP == 1 ? 1 : 2 * NP - 2;       /* barrier-separated phases */
    static constexpr int LP = L >= 32 ? ((L + L / 16) | 1) : L;
    static constexpr int SMEM_ELEMS = NP == 1 ? 1 : B * LP;    /* float2 units */
    static constexpr size_t SMEM_BYTES = sizeof(float2) * SMEM_ELEMS;
    static_assert(L * B >= 16 * 32 && (L * B) % (16 * 32) == 0, "tile must fill whole warps");
    static_assert(T <= 1024, "tile too large for one CTA");

    template <int P> struct Pass {
        static constexpr int R = fft_radix(L, P);
        static constexpr int NS = fft_ns(L, P);
        static constexpr int NBF = L / R;                     /* butterflies per signal */
        static constexpr int U = E / R;                       /* butterflies per thread */
        static constexpr bool BF = (P == 0 && LOAD_BF) || (P == NP - 1 && STORE_BF);
        static FDC_HD void map(int tid, int u, int& batch, int& j)
        {
            const int i = tid + u * T;
            if (BF) { batch = i % B; j = i / B; }
            else { batch = i / NBF; j = i % NBF; }
        }
    };
    /* physical smem slot of logical position pos of signal batch in exchange e (= written by pass e) */
    template <int EX> static FDC_HD int phys(int batch, int pos) { return batch * LP + pos + (EX == 0 ? (pos >> 4) : 0); }

    template <int P> static FDC_HD void twiddle_bfly(int tid, float2* v, const float2* tw)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            if (P > 0) {
                int batch, j; PS::map(tid, u, batch, j);
                const int k = j % PS::NS;
                const int step = k * (L / (PS::NS * PS::R));
#pragma unroll
                for (int t = 1; t < PS::R; t++) {
                    float2 w = fdc_ldg(tw + step * t);
                    if (DIR < 0) w.y = -w.y;
                    v[u * PS::R + t] = cmul(v[u * PS::R + t], w);
                }
            }
            Bfly<PS::R, DIR>::run(v + u * PS::R);
        }
    }
    template <int P, class Loader> static FDC_HD void load_global(int tid, float2* v, Loader& ld)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            const typename Loader::Ctx c = ld.begin(batch);       /* per-signal addressing, hoisted out of the element loop */
#pragma unroll
            for (int t = 0; t < PS::R; t++) v[u * PS::R + t] = ld.get(c, j + t * PS::NBF);
        }
    }
    template <int P> static FDC_HD void read_smem(int tid, float2* v, const float2* smem)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
#pragma unroll
            for (int t = 0; t < PS::R; t++) v[u * PS::R + t] = smem[phys<P - 1>(batch, j + t * PS::NBF)];
        }
    }
    template <int P> static FDC_HD void write_smem(int tid, const float2* v, float2* smem)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            const int k = j % PS::NS;
            const int o = (j - k) * PS::R + k;
#pragma unroll
            for (int t = 0; t < PS::R; t++) smem[phys<P>(batch, o + t * PS::NS)] = v[u * PS::R + t];
        }
    }
    template <int P, class Storer> static FDC_HD void store_global(int tid, const float2* v, Storer& st)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            const int k = j % PS::NS;
            const int o = (j - k) * PS::R + k;
            const typename Storer::Ctx c = st.begin(batch);
#pragma unroll
            for (int t = 0; t < PS::R; t++) st.put(c, o + t * PS::NS, v[u * PS::R + t]);
        }
    }

    /* one barrier-separated phase; v[16] is the calling thread's register tile and persists across phases */
    template <int PH, class Loader, class Storer>
    static FDC_HD void phase(int tid, float2* v, float2* smem, const float2* tw, Loader& ld, Storer& st)
    {
        if constexpr (NP == 1) {
            load_global<0>(tid, v, ld); twiddle_bfly<0>(tid, v, tw); store_global<0>(tid, v, st);
        } else if constexpr (PH == 0) {
            load_global<0>(tid, v, ld); twiddle_bfly<0>(tid, v, tw); write_smem<0>(tid, v, smem);
        } else if constexpr (PH % 2 == 1) {
            constexpr int P = (PH + 1) / 2;
            read_smem<P>(tid, v, smem); twiddle_bfly<P>(tid, v, tw);
            if constexpr (P == NP - 1) store_global<P>(tid, v, st);
        } else {
            constexpr int P = PH / 2;
            write_smem<P>(tid, v, smem);
        }
    }
};

#if defined(__CUDACC__)
/* device driver: run all phases of one tile with CTA barriers in between */
template <class ENG, int PH, class Loader, class Storer>
__device__ __forceinline__ void tile_fft_from(float2* v, float2* smem, const float2* tw, Loader& ld, Storer& st)
{
    ENG::template phase<PH>(threadIdx.x, v, smem, tw, ld, st);
    if constexpr (PH + 1 < ENG::NPH) {
        __syncthreads();
        tile_fft_from<ENG, PH + 1, Loader, Storer>(v, smem, tw, ld, st);
    }
}
template <class ENG, class Loader, class Storer>
__device__ __forceinline__ void tile_fft_run(float2* smem, const float2* tw, Loader& ld, Storer& st)
{
    float2 v[16];
    tile_fft_from<ENG, 0, Loader, Storer>(v, smem, tw, ld, st);
}
#endif

}  // namespace fdc
#endif
