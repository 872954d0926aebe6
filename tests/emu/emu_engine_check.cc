/* TEST INFRASTRUCTURE: host emulation of the tile-FFT phases and of the kernels' loader/storer functors
 * (index arithmetic check before GPU time is spent), see csrc/fdc_hd.h.  Every phase is stepped over all
 * thread ids of a CTA, phases in order = the barriers of the device code.  Checked against fp64 DFTs. */
#include "fdc_tile_fft.cuh"
#include "fdc_functors.cuh"
#include <vector>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <array>
#include <algorithm>
#include <cstring>
using namespace fdc;
typedef std::complex<double> cd;

static std::vector<float2> pass_twiddles(int L, int E = 16)
{
    std::vector<float2> h((size_t)fft_twsize(L, E), make_float2(1.f, 0.f));
    for (int p = 1; p < fft_npasses(L, E); p++) {
        const int R = fft_radix(L, p, E), NS = fft_ns(L, p, E), off = fft_twoff(L, p, E);
        for (int t = 1; t < R; t++) {
            if (!fft_twstored(t, R)) continue;
            for (int k = 0; k < NS; k++) {
                const double a = -2.0 * M_PI * (double)(((long)k * t) % ((long)NS * R)) / ((double)NS * R);
                h[(size_t)(off + fft_twrow(t, R) * NS + k)] = make_float2((float)cos(a), (float)sin(a));
            }
        }
    }
    return h;
}
/* fp64 reference FFT (iterative radix-2), sign = -1 forward, +1 backward */
static void fft64(std::vector<cd>& a, int sign)
{
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; k++) {
                const cd w = std::polar(1.0, sign * 2.0 * M_PI * (double)k / (double)len);
                const cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v; a[i + k + len / 2] = u - v;
            }
    }
}
static float frand() { return (float)rand() / RAND_MAX - 0.5f; }

/* run one tile through all phases */
template <class ENG, int PH, class LD, class ST> struct Run {
    static void go(std::vector<std::array<float2, 32>>& regs, float2* smem, const float2* tw, const LD& ld, const ST& st)
    {
        if (PH == 0) for (int tid = 0; tid < ENG::T; tid++) ENG::finish(tid, regs[tid].data(), ld);
        for (int tid = 0; tid < ENG::T; tid++) ENG::template phase<PH, false>(tid, regs[tid].data(), smem, tw, st);
        if constexpr (PH + 1 < ENG::NPH) Run<ENG, PH + 1, LD, ST>::go(regs, smem, tw, ld, st);
    }
};
template <class ENG, class Tiles> static void run_tiles(const Tiles& tiles, long ntiles, const float2* tw)
{
    std::vector<float2> smem(ENG::SMEM_ELEMS);
    std::vector<std::array<float2, 32>> regs(ENG::T);
    /* walk the tiles like a persistent CTA does: a stride that is not a multiple of ninner exercises tile_advance */
    const int ninner = tiles.ninner();
    const long stride = 3;
    for (long first = 0; first < stride && first < ntiles; first++) {
      TilePos pos = tile_split(first, ninner); const TilePos step = tile_split(stride, ninner);
      for (long tile = first; tile < ntiles; tile += stride, pos = tile_advance(pos, step, ninner)) {
        const TilePos chk = tile_split(tile, ninner);
        if (chk.inner != pos.inner || chk.outer != pos.outer) { printf("tile_advance mismatch at %ld\n", tile); exit(2); }
        auto ld = tiles.loader(pos); auto st = tiles.storer(pos);
        for (int tid = 0; tid < ENG::T; tid++) ENG::fetch(tid, regs[tid].data(), ld);
        Run<ENG, 0, decltype(ld), decltype(st)>::go(regs, smem.data(), tw, ld, st);
      }
    }
}
static double rel_err(const std::vector<cd>& want, const float2* got)
{
    double num = 0, den = 0;
    for (size_t i = 0; i < want.size(); i++) { num += std::norm(want[i] - cd(got[i].x, got[i].y)); den += std::norm(want[i]); }
    return sqrt(num / (den > 0 ? den : 1));
}
static int g_fail = 0;
static void report(const char* what, double e, double tol)
{
    printf("%-64s relL2=%.3e %s\n", what, e, e < tol ? "ok" : "FAIL");
    if (!(e < tol)) g_fail++;
}

/* ---- the plain engine at every length, both directions, with and without the shift permutations ---- */
template <int L, int DIR> static void check_plain(int shift)
{
    constexpr int B = L >= 4096 ? 1 : 4096 / L;
    typedef TileFFT<L, B, DIR, false, false> ENG;
    const long nvec = 2 * B + 1 < 3 ? 3 : B + 1;                       /* a ragged last tile */
    std::vector<float2> in((size_t)nvec * L), out((size_t)nvec * L, make_float2(7.f, 7.f));
    for (auto& v : in) { v.x = frand(); v.y = frand(); }
    const std::vector<float2> tw = pass_twiddles(L);
    PlainParams p; p.in = in.data(); p.out = out.data(); p.nvec = nvec; p.shift = shift;
    run_tiles<ENG>(PlainTiles<L, B, DIR>{p}, (nvec + B - 1) / B, tw.data());
    double worst = 0;
    for (long v = 0; v < nvec; v++) {
        std::vector<cd> a(L);
        for (int n = 0; n < L; n++) {
            const int src = (DIR < 0 && shift) ? (n + L / 2) % L : n;
            a[n] = cd(in[(size_t)v * L + src].x, in[(size_t)v * L + src].y);
        }
        if (L > 1) fft64(a, -DIR);
        std::vector<cd> want(L);
        for (int k = 0; k < L; k++) want[(DIR > 0 && shift) ? (k ^ (L / 2)) : k] = a[k];
        worst = std::max(worst, rel_err(want, out.data() + (size_t)v * L));
    }
    char name[128]; snprintf(name, sizeof name, "plain L=%d DIR=%d shift=%d (NP=%d, T=%d, smem=%zu)", L, DIR, shift, ENG::NP, ENG::T, ENG::SMEM_BYTES);
    report(name, worst, 2e-6);
}

/* ---- K1 small: overlap addressing, shift, scale ---- */
template <int N, int E = 16> static void check_fwd_small(int ovl, long nblocks)
{
    constexpr int B = N >= 4096 ? 1 : 4096 / N;
    typedef TileFFT<N, B, 1, false, false, E> ENG;
    const int hop = N - ovl;
    std::vector<float2> buf((size_t)ovl + (size_t)nblocks * hop), spec((size_t)nblocks * N);
    for (auto& v : buf) { v.x = frand(); v.y = frand(); }
    const std::vector<float2> tw = pass_twiddles(N, E);
    /* the ovl samples before in[0] come from the history buffer: what lies before in[0] in the stream buffer is poisoned */
    std::vector<float2> hist(buf.begin(), buf.begin() + ovl), stream(buf);
    for (int i = 0; i < ovl; i++) stream[(size_t)i] = make_float2(1e30f, -1e30f);
    FwdParams p; p.in = stream.data() + ovl; p.spec = spec.data(); p.nblocks = nblocks; p.hop = hop; p.ovl = ovl; p.N = N; p.scale = 1.0f / N; p.l2pf = 0;
    p.hist = hist.data(); p.head_blocks = ovl ? (ovl + hop - 1) / hop : 0; p.head_off = 0;
    run_tiles<ENG>(FwdTiles<N, B>{p}, (nblocks + B - 1) / B, tw.data());
    double worst = 0;
    for (long b = 0; b < nblocks; b++) {
        std::vector<cd> a(N);
        for (int n = 0; n < N; n++) a[n] = cd(buf[(size_t)b * hop + n].x, buf[(size_t)b * hop + n].y);
        fft64(a, -1);
        std::vector<cd> want(N);
        for (int k = 0; k < N; k++) want[k ^ (N / 2)] = a[k] / (double)N;
        worst = std::max(worst, rel_err(want, spec.data() + (size_t)b * N));
    }
    char name[128]; snprintf(name, sizeof name, "fwd_small N=%d E=%d ovl=%d nblocks=%ld", N, E, ovl, nblocks);
    report(name, worst, 2e-6);
}

/* ---- K1 big: four-step with column/row tiles ---- */
template <int N1, int N2> static void check_fwd_big(int ovl, long nblocks)
{
    constexpr int N = N1 * N2;
    constexpr int BC = N1 == 256 ? 16 : 4096 / N1, BR = N2 == 256 ? 16 : 4096 / N2;
    constexpr int EC = (N1 == 512 || N1 == 1024) ? 32 : 16, ER = (N2 == 512 || N2 == 1024) ? 32 : 16;    /* as fdc_k_fwd.cu big_points() */
    typedef TileFFT<N1, BC, 1, true, true, EC> CE;
    typedef TileFFT<N2, BR, 1, false, true, ER> RE;
    const int hop = N - ovl;
    std::vector<float2> buf((size_t)ovl + (size_t)nblocks * hop), mid((size_t)nblocks * N), spec((size_t)nblocks * N), tw4((size_t)N);
    for (auto& v : buf) { v.x = frand(); v.y = frand(); }
    for (long k1 = 0; k1 < N1; k1++)
        for (long n2 = 0; n2 < N2; n2++) {
            const double a = -2.0 * M_PI * (double)((k1 * n2) % N) / N;
            tw4[(size_t)(k1 * N2 + n2)] = make_float2((float)cos(a), (float)sin(a));
        }
    const std::vector<float2> twc = pass_twiddles(N1, EC), twr = pass_twiddles(N2, ER);
    std::vector<float2> hist(buf.begin(), buf.begin() + ovl), stream(buf);
    for (int i = 0; i < ovl; i++) stream[(size_t)i] = make_float2(1e30f, -1e30f);      /* must come from the history buffer */
    BigParams p; p.in = stream.data() + ovl; p.mid = mid.data(); p.spec = spec.data(); p.tw4 = tw4.data(); p.nblocks = nblocks;
    p.hop = hop; p.ovl = ovl; p.scale = 1.0f / N; p.hist = hist.data(); p.head_blocks = ovl ? (ovl + hop - 1) / hop : 0; p.head_off = 0;
    /* a persistent column CTA keeps one column tile and its twiddle slice; the emulator walks all tiles with "one CTA",
     * so the slice is rebuilt per column tile: run the tiles of one column tile at a time */
    {
        std::vector<float2> tws((size_t)N1 * BC);
        for (int ct = 0; ct < N2 / BC; ct++) {
            for (int tid = 0; tid < CE::T; tid++) CE::template last_pass_init<ColTwiddles<N1, N2, BC> >(tid, tws.data(), (const float2*)tw4.data(), ct);
            ColTiles<N1, N2, BC> tiles{p, tws.data()};
            std::vector<float2> smem(CE::SMEM_ELEMS);
            std::vector<std::array<float2, 32>> regs(CE::T);
            for (long b = 0; b < nblocks; b++) {
                TilePos pos; pos.inner = ct; pos.outer = (int)b;
                auto ld = tiles.loader(pos); auto st = tiles.storer(pos);
                for (int tid = 0; tid < CE::T; tid++) CE::fetch(tid, regs[tid].data(), ld);
                Run<CE, 0, decltype(ld), decltype(st)>::go(regs, smem.data(), twc.data(), ld, st);
            }
        }
    }
    run_tiles<RE>(RowTiles<N1, N2, BR>{p}, nblocks * (N1 / BR), twr.data());
    double worst = 0;
    for (long b = 0; b < nblocks; b++) {
        std::vector<cd> a(N);
        for (int n = 0; n < N; n++) a[n] = cd(buf[(size_t)b * hop + n].x, buf[(size_t)b * hop + n].y);
        fft64(a, -1);
        std::vector<cd> want(N);
        for (int k = 0; k < N; k++) want[k ^ (N / 2)] = a[k] / (double)N;
        worst = std::max(worst, rel_err(want, spec.data() + (size_t)b * N));
    }
    char name[128]; snprintf(name, sizeof name, "fwd_big N=%dx%d ovl=%d nblocks=%ld", N1, N2, ovl, nblocks);
    report(name, worst, 2e-6);
}

/* ---- K2: channel tiles, shared tables, phase selection, overlap discard, gain ---- */
template <int L, int E = 16, bool PACK = false> static void check_extract(int N, int nchan, long nb, int nphase)
{
    constexpr int B = E == 8 ? 2048 / L : (L >= 4096 ? 1 : 4096 / L);
    typedef TileFFT<L, B, -1, false, false, E> ENG;
    static_assert(ENG::E <= 32, "emulator register tile");
    std::vector<float2> spec((size_t)nb * N), tables((size_t)2 * nphase * L);
    for (auto& v : spec) { v.x = frand(); v.y = frand(); }
    for (auto& v : tables) { v.x = frand(); v.y = frand(); }
    std::vector<ChanDev> chans(nchan);
    long prefix = 0;
    const long call_blocks = nb + 3, call_blk0 = 2; const int glob_phase0 = 1 % nphase;
    for (int i = 0; i < nchan; i++) {
        ChanDev& c = chans[i];
        c.f = (int)(((long)i * (N - L)) / std::max(1, nchan - 1)); c.lout = L - L / 4 - (i % 3 == 1 ? 1 : 0); if (c.lout < 1) c.lout = 1;
        c.shift = (c.f + i) % nphase; c.tab_off = (i % 2) * (long)nphase * L; c.lout_prefix = prefix; c.gain = (float)(1 + i % 4); c.owner = 0; c.sink_prefix = 0; c.pad1 = 0;
        prefix += c.lout;
    }
    std::vector<float2> out((size_t)(call_blocks * prefix), make_float2(-9.f, -9.f));
    const std::vector<float2> tw = pass_twiddles(L, E);
    ExtractParams p; p.l2pf = 0; p.nsinks = 0; p.spec = spec.data(); p.spec_stride = N; p.tables = tables.data(); p.chans = chans.data();
    p.nsel = nchan; p.ny = (nchan + B - 1) / B; p.out = out.data(); p.nb = nb; p.call_blocks = call_blocks; p.call_blk0 = call_blk0;
    p.glob_phase0 = glob_phase0; p.nphase = nphase; p.tma_ok = 0; p.phase_mask = (nphase & (nphase - 1)) == 0 ? nphase - 1 : -1;
    p.bpt = 1;
    if (PACK) {              /* few channels: a tile holds all channels of bpt consecutive blocks */
        p.bpt = B / nchan; p.ny = 1;
        if (p.bpt < 2) { printf("packed case needs nchan <= B / 2\n"); exit(2); }
        run_tiles<ENG>(PackedExtractTiles<L, B>{p}, (nb + p.bpt - 1) / p.bpt, tw.data());
    } else
        run_tiles<ENG>(ExtractTiles<L, B>{p}, nb * p.ny, tw.data());
    double worst = 0;
    for (int i = 0; i < nchan; i++)
        for (long b = 0; b < nb; b++) {
            const ChanDev& c = chans[i];
            const int phase = (int)((((glob_phase0 + b) % nphase) * c.shift) % nphase);
            std::vector<cd> a(L);
            for (int n = 0; n < L; n++) {
                const int m = (n + L / 2) % L;
                const float2 x = spec[(size_t)b * N + c.f + m], w = tables[(size_t)(c.tab_off + (long)phase * L + m)];
                a[n] = cd(x.x, x.y) * cd(w.x, w.y);
            }
            if (L > 1) fft64(a, +1);
            std::vector<cd> want(c.lout);
            for (int k = 0; k < c.lout; k++) want[k] = a[L - c.lout + k] * (double)c.gain;
            worst = std::max(worst, rel_err(want, out.data() + (size_t)(call_blocks * c.lout_prefix + (call_blk0 + b) * c.lout)));
        }
    /* nothing outside the slabs' [call_blk0, call_blk0 + nb) rows may have been written */
    long stray = 0;
    for (int i = 0; i < nchan; i++)
        for (long b = 0; b < call_blocks; b++) {
            if (b >= call_blk0 && b < call_blk0 + nb) continue;
            const float2* r = out.data() + (size_t)(call_blocks * chans[i].lout_prefix + b * chans[i].lout);
            for (int k = 0; k < chans[i].lout; k++) if (r[k].x != -9.f) stray++;
        }
    char name[128]; snprintf(name, sizeof name, "extract%s L=%d E=%d N=%d nchan=%d nb=%ld nphase=%d stray=%ld", PACK ? " packed" : "", L, E, N, nchan, nb, nphase, stray);
    report(name, stray ? 1.0 : worst, 3e-6);
}

/* ---- activity-gated job list ---- */
template <int L> static void check_jobs(int N, int njobs)
{
    constexpr int B = L >= 4096 ? 1 : 4096 / L;
    typedef TileFFT<L, B, -1, false, false> ENG;
    const int rows = 3;
    std::vector<float2> spec((size_t)rows * N), hist((size_t)N), tables((size_t)4 * L);
    for (auto& v : spec) { v.x = frand(); v.y = frand(); }
    for (auto& v : hist) { v.x = frand(); v.y = frand(); }
    for (auto& v : tables) { v.x = frand(); v.y = frand(); }
    std::vector<ExtractJob> jobs(njobs);
    long off = 0;
    for (int i = 0; i < njobs; i++) {
        jobs[i].row = (i % 4) - 1; jobs[i].start = (i * 37) % (N - L + 1); jobs[i].tab_off = (i % 4) * L; jobs[i].skip = L / 4;
        jobs[i].dst_off = off; off += L - L / 4;
    }
    std::vector<float2> out((size_t)off);
    const std::vector<float2> tw = pass_twiddles(L);
    JobParams p; p.spec = spec.data(); p.spec_stride = N; p.hist = hist.data(); p.tables = tables.data(); p.jobs = jobs.data(); p.out = out.data(); p.njobs = njobs;
    run_tiles<ENG>(JobTiles<L, B>{p}, (njobs + B - 1) / B, tw.data());
    double worst = 0;
    for (int i = 0; i < njobs; i++) {
        const float2* src = (jobs[i].row < 0 ? hist.data() : spec.data() + (size_t)jobs[i].row * N) + jobs[i].start;
        std::vector<cd> a(L);
        for (int n = 0; n < L; n++) { const int m = (n + L / 2) % L; a[n] = cd(src[m].x, src[m].y) * cd(tables[jobs[i].tab_off + m].x, tables[jobs[i].tab_off + m].y); }
        fft64(a, +1);
        std::vector<cd> want(a.begin() + L / 4, a.end());
        worst = std::max(worst, rel_err(want, out.data() + jobs[i].dst_off));
    }
    char name[128]; snprintf(name, sizeof name, "jobs L=%d N=%d njobs=%d", L, N, njobs);
    report(name, worst, 3e-6);
}

int main()
{
#define PL(LL) check_plain<LL, 1>(1); check_plain<LL, -1>(1); check_plain<LL, 1>(0); check_plain<LL, -1>(0);
    PL(2) PL(4) PL(8) PL(16) PL(32) PL(64) PL(128) PL(256) PL(512) PL(1024) PL(2048) PL(4096) PL(8192) PL(16384)
#undef PL
    check_fwd_small<16>(4, 300); check_fwd_small<64>(16, 70); check_fwd_small<1024>(512, 9); check_fwd_small<4096>(1024, 3);
    check_fwd_small<8192>(2048, 2); check_fwd_small<16384>(2048, 2); check_fwd_small<2048>(1536, 5);
    check_fwd_small<512, 32>(128, 20); check_fwd_small<1024, 32>(512, 9);
    check_fwd_small<8192, 32>(2048, 2); check_fwd_small<16384, 32>(4096, 2);
    check_fwd_big<64, 64>(1024, 3); check_fwd_big<64, 128>(2048, 3); check_fwd_big<128, 128>(4096, 2);
    check_fwd_big<128, 256>(8192, 2); check_fwd_big<256, 256>(16384, 2); check_fwd_big<256, 512>(32768, 1);
    check_fwd_big<512, 512>(65536, 1); check_fwd_big<512, 1024>(131072, 1); check_fwd_big<256, 256>(49152, 3);
    check_extract<2>(64, 5, 3, 4); check_extract<8>(64, 7, 3, 4); check_extract<16>(256, 300, 2, 3); check_extract<64>(1024, 16, 5, 2);
    check_extract<128>(4096, 70, 3, 4); check_extract<256>(8192, 64, 3, 4); check_extract<512>(8192, 19, 5, 4);
    check_extract<1024>(4096, 5, 3, 4); check_extract<4096>(16384, 3, 2, 4); check_extract<8192>(16384, 2, 2, 8);
    check_extract<64, 8>(1024, 16, 5, 2); check_extract<128, 8>(4096, 70, 3, 4); check_extract<256, 8>(8192, 64, 3, 4);
    check_extract<512, 8>(8192, 19, 5, 4); check_extract<1024, 8>(4096, 5, 3, 4); check_extract<2048, 8>(8192, 3, 2, 4);
    check_extract<512, 32>(8192, 19, 5, 4); check_extract<1024, 32>(4096, 5, 3, 4); check_extract<512, 32>(65536, 256, 2, 4);
    check_extract<64, 16, true>(1024, 16, 11, 2); check_extract<64, 8, true>(1024, 16, 7, 4); check_extract<256, 16, true>(4096, 3, 17, 4);
    check_extract<128, 8, true>(2048, 5, 9, 3); check_extract<1024, 32, true>(4096, 1, 9, 4); check_extract<512, 32, true>(8192, 3, 6, 4);
    check_extract<2, 16, true>(64, 5, 700, 4);
    check_jobs<64>(1024, 70); check_jobs<512>(4096, 11); check_jobs<16>(256, 300);
    printf("%s\n", g_fail ? "EMU FAILED" : "EMU OK");
    return g_fail ? 1 : 0;
}
