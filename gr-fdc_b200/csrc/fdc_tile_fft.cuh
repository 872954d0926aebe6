/* fdc_tile_fft.cuh -- CTA-level batched complex-fp32 FFT ("tile FFT").
 *
 * One CTA transforms a tile of B independent length-L signals (B*L = E*T elements, T threads,
 * E = 16 (or 8) elements per thread held in registers).  Stockham autosort passes with register radix-16/8/4/2
 * butterflies; between passes the tile is exchanged through shared memory (in place: read all ->
 * barrier -> write all -> barrier).  The first pass takes its operands from global memory through a
 * Loader functor (overlap-save addressing, table multiply, half swap ... are fused there), the last
 * pass hands its results to a Storer functor (fft-shift, four-step twiddle, overlap discard, gain).
 *
 * Pass p (radix R, Ns = product of the earlier radices), butterfly j in [0, L/R):
 *      k    = j mod Ns
 *      in   : x[j + t*L/R] * W_{Ns*R}^{k t},  t = 0..R-1
 *      out  : y[(j-k)*R + k + t*Ns]
 * Every access of a butterfly is "base + t * compile-time stride": the functors and the shared-memory
 * exchange compute one address per butterfly and the 16 element accesses become immediate offsets.
 * The twiddles of pass p are stored as a [row(t)][k] table (k fastest) so that lanes with consecutive k
 * read consecutive entries.
 *
 * The first exchange (Ns = 1, a thread writes 16 consecutive points) is stored with one pad slot per
 * 16 points (pos + pos/16); Ns >= 16 afterwards and the later exchanges are conflict free as is.  The
 * per-signal stride LP is odd so that the "batch-fast" thread mappings (lanes walk over signals; used
 * when the global side is a strided column tile) stay conflict free too.
 *
 * Global loads are split in two steps, fetch (raw LDGs of the NEXT tile, issued before the current
 * tile is computed so that they are in flight during the butterflies) and finish (the functor's
 * arithmetic on the fetched values).  The code is organised in PHASES separated by CTA barriers so
 * that the same functions run on the device and under the host emulator (see fdc_hd.h). */
#ifndef FDC_TILE_FFT_CUH
#define FDC_TILE_FFT_CUH
#include <type_traits>
#include "fdc_bfly.cuh"
#include "fdc_tma.cuh"
#include "fdc_functors.cuh"

namespace fdc {

/* Radix plan of a length-L transform with E points per thread: the first pass has radix E (the thread's whole register
 * tile is one butterfly), the rest of L/E is split into radices <= E with the larger radix last (fewer store contexts per
 * thread in the pass that talks to global memory). */
constexpr int fft_next_radix(int rem, int E)
{
    if (rem >= 64 || rem == E) return E;
    if (rem == 32) return 4;                            /* 32 = 4 * 8 */
    if (rem == 16) return 4;                            /* only reached for E = 8: 16 = 4 * 4 */
    return rem;
}
/* E = 32 saves one pass (= one shared-memory exchange) where the length allows it:
 * 512 = 16 * 32, 1024 = 32 * 32 (2 passes instead of 3), 8192 = 32 * 16 * 16, 16384 = 32 * 16 * 32 (3 instead of 4) */
constexpr int fft_radix32(int L, int p)
{
    return L == 512 ? (p == 0 ? 16 : 32) : L == 1024 ? 32 : L == 8192 ? (p == 0 ? 32 : 16) : (p == 1 ? 16 : 32);
}
constexpr int fft_npasses(int L, int E = 16)
{
    if (E == 32) return L <= 1024 ? 2 : 3;
    if (L <= E) return 1;
    int n = 1, rem = L / E;
    while (rem > 1) { rem /= fft_next_radix(rem, E); n++; }
    return n;
}
constexpr int fft_radix(int L, int p, int E = 16)
{
    if (E == 32) return fft_radix32(L, p);
    if (L <= E) return L;
    if (p == 0) return E;
    int rem = L / E, q = 1, r = 1;
    while (true) {
        r = fft_next_radix(rem, E);
        if (q == p) return r;
        rem /= r; q++;
    }
}
constexpr int fft_ns(int L, int p, int E = 16)
{
    int ns = 1;
    for (int q = 0; q < p; q++) ns *= fft_radix(L, q, E);
    return ns;
}
/* FDC_TWGEN: only the rows t = 1 and t = 8, 16, 24 of a pass's twiddle table exist (the powers in between are generated,
 * twiddle_bfly); passes of radix < 8 keep all their rows.  fft_twrow(t, R) is the row of W^{k t}. */
#if defined(FDC_TWGEN)
constexpr bool fft_twgen(int R) { return R >= 8; }
#else
constexpr bool fft_twgen(int) { return false; }
#endif
constexpr int fft_twrows(int R) { return fft_twgen(R) ? 1 + (R - 1) / 8 : R - 1; }
constexpr int fft_twrow(int t, int R) { return fft_twgen(R) ? (t == 1 ? 0 : t / 8) : t - 1; }
constexpr bool fft_twstored(int t, int R) { return !fft_twgen(R) || t == 1 || t % 8 == 0; }
/* offset of pass p's twiddles inside the per-length table, and the table's size (float2 units) */
constexpr int fft_twoff(int L, int p, int E = 16)
{
    int off = 0;
    for (int q = 1; q < p; q++) off += fft_twrows(fft_radix(L, q, E)) * fft_ns(L, q, E);
    return off;
}
constexpr int fft_twsize(int L, int E = 16) { const int n = fft_twoff(L, fft_npasses(L, E), E); return n < 1 ? 1 : n; }
constexpr int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

/* storers may offer two code paths for a whole butterfly (e.g. unit gain): detected by a HAS_VARIANT member */
template <class S, class = void> struct storer_has_variant { static constexpr bool value = false; };
template <class S> struct storer_has_variant<S, decltype((void)S::HAS_VARIANT)> { static constexpr bool value = S::HAS_VARIANT; };

/* loaders whose first tiles read through a second segment (the overlap-save history): HAS_HEAD, is_head(), fetch_head() */
template <class Ld, class = void> struct loader_has_head { static constexpr bool value = false; };
template <class Ld> struct loader_has_head<Ld, decltype((void)Ld::HAS_HEAD)> { static constexpr bool value = Ld::HAS_HEAD; };

template <int L_, int B_, int DIR, bool LOAD_BF, bool STORE_BF, int E_ = 16>
struct TileFFT {
    static constexpr int L = L_, B = B_, E = E_;               /* E points per thread (16, or 8 for twice the warps) */
    static constexpr int T = L * B / E;                       /* threads per CTA */
    static constexpr int NP = fft_npasses(L, E);
    static constexpr int NPH = NP == 1 ? 1 : 2 * NP - 2;       /* barrier-separated phases */
    /* exchange 0 is padded by one slot per PADW points.  16 (32 for a radix-32 first pass) keeps both the first pass's
     * writes (a thread stores R0 <= PADW consecutive points) and the second pass's reads (lanes walk consecutive points)
     * on 32 distinct banks per 16-lane phase of the 64-bit accesses; the per-signal stride is odd only for the batch-fast
     * mappings, where lanes walk over signals (bank model: tests/emu/bank_model.py) */
    static constexpr int PADW = fft_radix(L, 0, E) > 16 ? 32 : 16;
    static constexpr int LP = NP > 1 ? ((LOAD_BF || STORE_BF) ? ((L + L / PADW) | 1) : (L + L / PADW)) : L;
    static constexpr int SMEM_ELEMS = NP == 1 ? 1 : B * LP;    /* float2 units */
    static constexpr size_t SMEM_BYTES = sizeof(float2) * SMEM_ELEMS;
    static constexpr int TWSIZE = fft_twsize(L, E);
    static_assert(E == 16 || E == 8 || (E == 32 && (L == 512 || L == 1024 || L == 8192 || L == 16384)), "8, 16 or (selected lengths) 32 points per thread");
    static_assert(L * B >= E * 32 && (L * B) % (E * 32) == 0, "tile must fill whole warps");
    static_assert(T <= 1024, "tile too large for one CTA");

    template <int P> struct Pass {
        static constexpr int R = fft_radix(L, P, E);
        static constexpr int NS = fft_ns(L, P, E);
        static constexpr int NBF = L / R;                     /* butterflies per signal */
        static constexpr int U = E / R;                       /* butterflies per thread */
        static constexpr int TWOFF = fft_twoff(L, P, E);
        static constexpr bool BF = (P == 0 && LOAD_BF) || (P == NP - 1 && STORE_BF);
        static FDC_HD void map(int tid, int u, int& batch, int& j)
        {
            const int i = tid + u * T;
            if (BF) { batch = i % B; j = i / B; }
            else { batch = i / NBF; j = i % NBF; }
        }
    };
    /* physical smem slot of logical position pos of signal batch in exchange e (= written by pass e) */
    template <int EX> static FDC_HD int phys(int batch, int pos) { return batch * LP + pos + (EX == 0 ? (pos >> ilog2(PADW)) : 0); }
    /* the same for pos + c with a compile-time c: returns the constant to add to phys(batch, pos) (c % PADW == 0 or EX > 0,
     * or pos % PADW == 0 and c < PADW) */
    template <int EX> static constexpr int phys_step(int c) { return c + (EX == 0 ? c / PADW : 0); }

    /* ---- global side --------------------------------------------------------------------------------------- */
    template <class Loader> static FDC_HD void fetch(int tid, float2* raw, const Loader& ld)
    {
        typedef Pass<0> PS;
        if constexpr (loader_has_head<Loader>::value) {
            if (ld.is_head()) {                             /* tile-uniform: only the first tile(s) of a call */
#pragma unroll
                for (int u = 0; u < PS::U; u++) {
                    int batch, j; PS::map(tid, u, batch, j);
                    const typename Loader::Ctx c = ld.begin(batch, j);
#pragma unroll
                    for (int t = 0; t < PS::R; t++) raw[u * PS::R + t] = ld.template fetch_head<PS::R, PS::NBF>(c, t);
                }
                return;
            }
        }
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            const typename Loader::Ctx c = ld.begin(batch, j);
#pragma unroll
            for (int t = 0; t < PS::R; t++) raw[u * PS::R + t] = ld.template fetch<PS::R, PS::NBF>(c, t);
        }
    }
    template <class Loader> static FDC_HD void finish(int tid, float2* v, const Loader& ld)
    {
        typedef Pass<0> PS;
        if (!Loader::HAS_FINISH) return;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            const typename Loader::Ctx c = ld.begin(batch, j);
#pragma unroll
            for (int t = 0; t < PS::R; t++) v[u * PS::R + t] = ld.template finish<PS::R, PS::NBF>(c, t, v[u * PS::R + t]);
        }
    }
    template <int P, class Storer> static FDC_HD void store_global(int tid, const float2* v, const Storer& st)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            /* last pass: NS * R == L, hence k == j and the outputs are j + t*NS */
            const typename Storer::Ctx c = st.begin(batch, j);
            if constexpr (storer_has_variant<Storer>::value) {
                if (st.variant(c)) {
#pragma unroll
                    for (int t = 0; t < PS::R; t++) st.template put<PS::R, PS::NS, true>(c, t, v[u * PS::R + t]);
                } else {
#pragma unroll
                    for (int t = 0; t < PS::R; t++) st.template put<PS::R, PS::NS, false>(c, t, v[u * PS::R + t]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < PS::R; t++) st.template put<PS::R, PS::NS>(c, t, v[u * PS::R + t]);
            }
        }
    }

    /* per-CTA constants of the last pass (e.g. the four-step twiddles): Init::init_one<R, NS>(args..., batch, o) for every
     * butterfly this thread owns in the last pass */
    template <class Init, class... A> static FDC_HD void last_pass_init(int tid, A... args)
    {
        typedef Pass<NP - 1> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            Init::template init_one<PS::R, PS::NS>(args..., batch, j);
        }
    }

    /* ---- butterflies ---------------------------------------------------------------------------------------- */
    /* TWS: tw points to a shared-memory copy of the table (plain loads); otherwise read-only global loads */
    /* Pass twiddles.  FDC_TWGEN: only W^k and every 8th power come from the table, the powers in between are generated with one
     * complex multiply each (cur *= W^k): two packed instructions instead of a 64-bit shared-memory load per element.  The
     * kernels are bound by the L1 / shared-memory data path (238 B per input sample go through it on cfg4), not by
     * the FP32 pipe; anchors every 8 powers keep the accumulated rounding below 7 ulp of the twiddle. */
    template <int P, bool TWS> static FDC_HD void twiddle_bfly(int tid, float2* v, const float2* tw)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            if (P > 0) {
                int batch, j; PS::map(tid, u, batch, j);
                const float2* twp = tw + PS::TWOFF + (j % PS::NS);
                float2 w1 = make_float2(1.f, 0.f), cur = w1;
#pragma unroll
                for (int t = 1; t < PS::R; t++) {
                    float2 w;
                    if (fft_twstored(t, PS::R)) {
                        w = TWS ? twp[fft_twrow(t, PS::R) * PS::NS] : fdc_ldg(twp + fft_twrow(t, PS::R) * PS::NS);
                        if (t == 1) w1 = w;
                    } else w = cmul(cur, w1);
                    cur = w;
                    const float2 a = v[u * PS::R + t];
                    v[u * PS::R + t] = DIR > 0 ? cmul(a, w) : cmulc(a, w);
                }
            }
            Bfly<PS::R, DIR>::run(v + u * PS::R);
        }
    }

    /* ---- shared-memory exchange ----------------------------------------------------------------------------- */
    template <int P> static FDC_HD void read_smem(int tid, float2* v, const float2* smem)
    {
        typedef Pass<P> PS;
        static_assert(P > 1 || PS::NBF % PADW == 0 || 2 * PS::NBF == PADW, "reads of the padded first exchange need a PADW-aligned butterfly stride");
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            const float2* s = smem + phys<P - 1>(batch, j);
#pragma unroll
            for (int t = 0; t < PS::R; t++) v[u * PS::R + t] = s[phys_step<P - 1>(t * PS::NBF)];
        }
    }
    template <int P> static FDC_HD void write_smem(int tid, const float2* v, float2* smem)
    {
        typedef Pass<P> PS;
#pragma unroll
        for (int u = 0; u < PS::U; u++) {
            int batch, j; PS::map(tid, u, batch, j);
            const int k = j % PS::NS;
            const int o = (j - k) * PS::R + k;               /* P == 0: o = 16 j, offsets t < 16; P > 0: offsets t*NS, NS % 16 == 0 */
            float2* s = smem + phys<P>(batch, o);
#pragma unroll
            for (int t = 0; t < PS::R; t++) s[phys_step<P>(t * PS::NS)] = v[u * PS::R + t];
        }
    }

    /* one barrier-separated phase; v[16] is the calling thread's register tile and persists across phases.
     * Phase 0 expects the finished (fetch + finish) values in v. */
    template <int PH, bool TWS, class Storer>
    static FDC_HD void phase(int tid, float2* v, float2* smem, const float2* tw, const Storer& st)
    {
        if constexpr (NP == 1) {
            twiddle_bfly<0, TWS>(tid, v, tw); store_global<0>(tid, v, st);
        } else if constexpr (PH == 0) {
            twiddle_bfly<0, TWS>(tid, v, tw); write_smem<0>(tid, v, smem);
        } else if constexpr (PH % 2 == 1) {
            constexpr int P = (PH + 1) / 2;
            read_smem<P>(tid, v, smem); twiddle_bfly<P, TWS>(tid, v, tw);
            if constexpr (P == NP - 1) store_global<P>(tid, v, st);
        } else {
            constexpr int P = PH / 2;
            write_smem<P>(tid, v, smem);
        }
    }
};

#if defined(__CUDACC__)
/* twiddles are served from shared memory when the per-length table is small (<= 8 KB) */
constexpr bool tw_in_smem(int L, int E = 16) { return fft_npasses(L, E) > 1 && fft_twsize(L, E) <= 1024; }
constexpr int tw_smem_elems(int L, int E = 16) { return tw_in_smem(L, E) ? fft_twsize(L, E) : 0; }

/* device driver: run all phases of one tile with CTA barriers in between (v holds the finished first-pass inputs) */
template <class ENG, int PH, bool TWS, class Storer>
__device__ __forceinline__ void tile_fft_from(float2* v, float2* smem, const float2* tw, const Storer& st)
{
    ENG::template phase<PH, TWS>(threadIdx.x, v, smem, tw, st);
    if constexpr (PH + 1 < ENG::NPH) {
        __syncthreads();
        tile_fft_from<ENG, PH + 1, TWS, Storer>(v, smem, tw, st);
    }
}
/* Persistent tile loop: the CTA walks tiles first, first + stride, ... < ntiles; with PF the global loads of tile
 * i+1 are issued before tile i is computed (register prefetch).  Tiles::loader(pos) / Tiles::storer(pos) make the
 * functors from the (inner, outer) split of the tile index, which is advanced without divisions.
 * Order inside an iteration: finish (the loader's own table loads + arithmetic) -> prefetch of the next tile ->
 * butterflies.  The only global loads in flight during the butterflies are the prefetch: the twiddles come from
 * shared memory, so no instruction waits on a scoreboard that it shares with a DRAM access. */
template <class T, class = void> struct tiles_l2_prefetch { static constexpr bool value = false; };
template <class T> struct tiles_l2_prefetch<T, typename std::enable_if<T::HAS_L2_PREFETCH>::type> { static constexpr bool value = true; };
template <class ENG, bool PF, class Tiles, bool TW_SMEM = true>
__device__ __forceinline__ void tile_fft_loop(float2* smem, const float2* tw_global, const Tiles& tiles, long first, long stride, long ntiles)
{
    constexpr bool TWS = TW_SMEM && tw_in_smem(ENG::L, ENG::E);       /* TW_SMEM = false: twiddles stay in global memory / L1 (saves 4 KB per CTA) */
    const float2* tw = tw_global;
    /* programmatic dependent launch: let the next kernel of the stream start its prologue now, run ours (constant tables
     * only), then wait until the previous kernel's results are complete and visible */
    cudaTriggerProgrammaticLaunchCompletion();
    if constexpr (TWS) {
        float2* tws = smem + ENG::SMEM_ELEMS;
        for (int i = threadIdx.x; i < ENG::TWSIZE; i += ENG::T) tws[i] = tw_global[i];
        tw = tws;
        __syncthreads();
    }
    cudaGridDependencySynchronize();
    if (first >= ntiles) return;
    const int ninner = tiles.ninner();
    const TilePos step = tile_split(stride, ninner);
    TilePos pos = tile_split(first, ninner);
    float2 v[ENG::E];
    if (PF) ENG::fetch(threadIdx.x, v, tiles.loader(pos));
    for (long tile = first;;) {
        const long next = tile + stride;
        const TilePos npos = tile_advance(pos, step, ninner);
        if constexpr (PF) {
            float2 nx[ENG::E];
            ENG::finish(threadIdx.x, v, tiles.loader(pos));
            asm volatile("" ::: "memory");
            if (next < ntiles) ENG::fetch(threadIdx.x, nx, tiles.loader(npos));
            asm volatile("" ::: "memory");
            tile_fft_from<ENG, 0, TWS>(v, smem, tw, tiles.storer(pos));
            if (next >= ntiles) break;
#pragma unroll
            for (int i = 0; i < ENG::E; i++) v[i] = nx[i];
        } else {
            ENG::fetch(threadIdx.x, v, tiles.loader(pos));
            if constexpr (tiles_l2_prefetch<Tiles>::value) { if (next < ntiles) tiles.prefetch_l2(npos, (int)threadIdx.x); }
            ENG::finish(threadIdx.x, v, tiles.loader(pos));
            tile_fft_from<ENG, 0, TWS>(v, smem, tw, tiles.storer(pos));
            if (next >= ntiles) break;
        }
        tile = next; pos = npos;
        if (ENG::NP > 1) __syncthreads();          /* the exchange buffer is reused by the next tile */
    }
}
#endif

}  // namespace fdc
#endif
