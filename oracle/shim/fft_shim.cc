/* Oracle shim (TEST INFRASTRUCTURE ONLY) -- FFT behind gr::fft::fft_complex.
 *
 * Restates what FFTW3f computes for fftwf_plan_dft_1d: out[k] = sum_n in[n] * exp(-/+ j 2 pi k n / N),
 * sign - for forward, + for backward, no normalisation (FFTW manual, "What FFTW Really Computes").
 * FFTW's exact operation order is plan dependent and unpinned, so no fp32 bit pattern is pinned here;
 * mode 0 computes in fp64 and rounds once (error <= 0.5 ulp of fp32 per output), which bounds any
 * correct fp32 implementation to ~1e-7 relative.  Power-of-two sizes only (all gr-FDC uses). */
#include <gnuradio/fft/fft.h>
#include <vector>
#include <map>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <memory>

static int g_fft_mode = 0;
extern "C" void fdc_shim_set_fft_mode(int mode) { g_fft_mode = mode; }
extern "C" int fdc_shim_get_fft_mode(void) { return g_fft_mode; }

namespace {
typedef std::complex<double> cd;
typedef std::complex<float> cf;

struct plan64 { int n, logn; std::vector<cd> tw; std::vector<int> rev; };
struct plan32 { int n; std::vector<cf> tw; };

std::mutex g_mutex;
std::map<int, std::shared_ptr<plan64> > g_p64;
std::map<int, std::shared_ptr<plan32> > g_p32;

std::shared_ptr<plan64> get_plan64(int n)
{
    std::lock_guard<std::mutex> g(g_mutex);
    std::shared_ptr<plan64>& p = g_p64[n];
    if (!p) {
        p.reset(new plan64);
        p->n = n; p->logn = 0; while ((1 << p->logn) < n) p->logn++;
        if ((1 << p->logn) != n) throw std::invalid_argument("fft shim: size must be a power of two");
        p->tw.resize(n / 2 > 0 ? n / 2 : 1);
        for (int k = 0; k < n / 2; k++) {
            const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)n;
            p->tw[k] = cd((double)cosl(a), (double)sinl(a));
        }
        p->rev.resize(n);
        for (int i = 0; i < n; i++) { int r = 0; for (int b = 0; b < p->logn; b++) if (i & (1 << b)) r |= 1 << (p->logn - 1 - b); p->rev[i] = r; }
    }
    return p;
}
std::shared_ptr<plan32> get_plan32(int n)
{
    std::lock_guard<std::mutex> g(g_mutex);
    std::shared_ptr<plan32>& p = g_p32[n];
    if (!p) {
        p.reset(new plan32); p->n = n;
        p->tw.resize(n);
        for (int k = 0; k < n; k++) { const double a = -2.0 * M_PI * (double)k / (double)n; p->tw[k] = cf((float)cos(a), (float)sin(a)); }
    }
    return p;
}

/* fp64 iterative radix-2 DIT, bit-reversed load */
void fft64(const plan64& p, bool fwd, const cf* in, cf* out)
{
    const int n = p.n;
    std::vector<cd> a(n);
    for (int i = 0; i < n; i++) a[p.rev[i]] = cd(in[i].real(), in[i].imag());
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; k++) {
                cd w = p.tw[k * step]; if (!fwd) w = std::conj(w);
                const cd u = a[i + k], v = cd(a[i + k + half].real() * w.real() - a[i + k + half].imag() * w.imag(),
                                              a[i + k + half].real() * w.imag() + a[i + k + half].imag() * w.real());
                a[i + k] = u + v; a[i + k + half] = u - v;
            }
    }
    for (int i = 0; i < n; i++) out[i] = cf((float)a[i].real(), (float)a[i].imag());
}

/* fp32 Stockham autosort, radix-4 passes (+ one radix-2 pass when log2 n is odd).  Ping-pongs x <-> y. */
void fft32(const plan32& p, bool fwd, const cf* in, cf* out)
{
    const int n = p.n;
    if (n == 1) { out[0] = in[0]; return; }
    std::vector<cf> buf0(in, in + n), buf1(n);
    cf* x = buf0.data(); cf* y = buf1.data();
    const cf* tw = p.tw.data();
    const float sgn = fwd ? 1.0f : -1.0f;       /* conj twiddles + swap +-j for backward */
    int ns = 1;
    int logn = 0; while ((1 << logn) < n) logn++;
    if (logn & 1) {                              /* radix-2 first (ns = 1, trivial twiddles) */
        const int h = n / 2;
        for (int j = 0; j < h; j++) { const cf a = x[j], b = x[j + h]; y[2 * j] = a + b; y[2 * j + 1] = a - b; }
        std::swap(x, y); ns = 2;
    }
    while (ns < n) {
        const int q = n / 4, tstep = n / (ns * 4);
        for (int j0 = 0; j0 < q; j0 += ns) {
            for (int k = 0; k < ns; k++) {
                const int j = j0 + k;
                cf w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep];
                if (!fwd) { w1 = std::conj(w1); w2 = std::conj(w2); w3 = std::conj(w3); }
                const cf a = x[j];
                const cf b0 = x[j + q], c0 = x[j + 2 * q], d0 = x[j + 3 * q];
                const cf b(b0.real() * w1.real() - b0.imag() * w1.imag(), b0.real() * w1.imag() + b0.imag() * w1.real());
                const cf c(c0.real() * w2.real() - c0.imag() * w2.imag(), c0.real() * w2.imag() + c0.imag() * w2.real());
                const cf d(d0.real() * w3.real() - d0.imag() * w3.imag(), d0.real() * w3.imag() + d0.imag() * w3.real());
                const cf s0 = a + c, s1 = a - c, s2 = b + d, s3 = b - d;
                const cf js3(sgn * s3.imag(), -sgn * s3.real());       /* -j*s3 (fwd) / +j*s3 (bwd) */
                const int o = (j0 * 4) + k;                            /* (j/ns)*ns*4 + k */
                y[o] = s0 + s2; y[o + ns] = s1 + js3; y[o + 2 * ns] = s0 - s2; y[o + 3 * ns] = s1 - js3;
            }
        }
        std::swap(x, y); ns *= 4;
    }
    memcpy(out, x, sizeof(cf) * n);
}
}  // namespace

namespace gr { namespace fft {
void fft_exec(int n, bool forward, const gr_complex* in, gr_complex* out)
{
    if (g_fft_mode == 1) fft32(*get_plan32(n), forward, in, out);
    else fft64(*get_plan64(n), forward, in, out);
}
fft_complex::fft_complex(int fft_size, bool forward, int) : d_size(fft_size), d_forward(forward)
{
    if (fft_size < 1 || (fft_size & (fft_size - 1))) throw std::invalid_argument("fft shim: size must be a power of two");
    void* a = 0; void* b = 0;
    if (posix_memalign(&a, 64, sizeof(gr_complex) * fft_size) || posix_memalign(&b, 64, sizeof(gr_complex) * fft_size))
        throw std::runtime_error("fft shim: alloc");
    d_in = (gr_complex*)a; d_out = (gr_complex*)b;
}
fft_complex::~fft_complex() { free(d_in); free(d_out); }
void fft_complex::execute() { fft_exec(d_size, d_forward, d_in, d_out); }
}}
