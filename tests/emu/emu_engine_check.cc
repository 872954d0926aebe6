/* TEST INFRASTRUCTURE: host emulation of the tile-FFT phases (index arithmetic check), see csrc/fdc_hd.h */
#include "fdc_tile_fft.cuh"
#include <vector>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <array>
using namespace fdc;
typedef std::complex<double> cd;

struct Ld { typedef const float2* Ctx; const float2* in; int L; Ctx begin(int batch) const { return in + batch * L; } float2 get(const Ctx& c, int n) const { return c[n]; } };
struct St { typedef float2* Ctx; float2* out; int L; Ctx begin(int batch) const { return out + batch * L; } void put(const Ctx& c, int k, float2 v) const { c[k] = v; } };

template <class ENG, int PH> struct Run {
    static void go(std::vector<std::array<float2, 16>>& regs, float2* smem, const float2* tw, Ld& ld, St& st)
    {
        for (int tid = 0; tid < ENG::T; tid++) ENG::template phase<PH>(tid, regs[tid].data(), smem, tw, ld, st);
        if constexpr (PH + 1 < ENG::NPH) Run<ENG, PH + 1>::go(regs, smem, tw, ld, st);
    }
};

template <int L, int B, int DIR, bool LBF, bool SBF> double check()
{
    typedef TileFFT<L, B, DIR, LBF, SBF> ENG;
    std::vector<float2> in(L * B), out(L * B), smem(ENG::SMEM_ELEMS), tw(L);
    for (auto& v : in) { v.x = (float)rand() / RAND_MAX - 0.5f; v.y = (float)rand() / RAND_MAX - 0.5f; }
    for (int m = 0; m < L; m++) { tw[m].x = (float)cos(-2.0 * M_PI * m / L); tw[m].y = (float)sin(-2.0 * M_PI * m / L); }
    std::vector<std::array<float2, 16>> regs(ENG::T);
    Ld ld{in.data(), L}; St st{out.data(), L};
    Run<ENG, 0>::go(regs, smem.data(), tw.data(), ld, st);
    double num = 0, den = 0;
    for (int b = 0; b < B; b++)
        for (int k = 0; k < L; k++) {
            cd acc = 0;
            for (int n = 0; n < L; n++) acc += cd(in[b * L + n].x, in[b * L + n].y) * std::polar(1.0, -DIR * 2.0 * M_PI * (double)((long)k * n % L) / L);
            cd d = acc - cd(out[b * L + k].x, out[b * L + k].y);
            num += std::norm(d); den += std::norm(acc);
        }
    double e = sqrt(num / den);
    printf("L=%5d B=%3d DIR=%2d LBF=%d SBF=%d NP=%d T=%4d smem=%6zu  relL2=%.3e %s\n", L, B, DIR, LBF, SBF, ENG::NP, ENG::T, ENG::SMEM_BYTES, e, e < 2e-6 ? "ok" : "FAIL");
    return e;
}
int main()
{
    double w = 0;
    w = std::max(w, check<2, 256, 1, false, false>());
    w = std::max(w, check<4, 128, -1, false, false>());
    w = std::max(w, check<8, 64, -1, false, false>());
    w = std::max(w, check<16, 32, 1, false, false>());
    w = std::max(w, check<32, 16, -1, false, false>());
    w = std::max(w, check<64, 16, -1, false, false>());
    w = std::max(w, check<128, 16, -1, false, false>());
    w = std::max(w, check<256, 16, -1, false, false>());
    w = std::max(w, check<256, 16, 1, true, true>());
    w = std::max(w, check<256, 16, 1, false, true>());
    w = std::max(w, check<512, 8, -1, false, false>());
    w = std::max(w, check<512, 16, 1, true, true>());
    w = std::max(w, check<1024, 4, 1, false, false>());
    w = std::max(w, check<1024, 4, -1, false, false>());
    w = std::max(w, check<2048, 2, 1, false, false>());
    w = std::max(w, check<4096, 1, 1, false, false>());
    w = std::max(w, check<4096, 1, -1, false, false>());
    w = std::max(w, check<8192, 1, 1, false, false>());
    w = std::max(w, check<16384, 1, 1, false, false>());
    return w < 2e-6 ? 0 : 1;
}
