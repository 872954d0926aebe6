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
struct plan32 { int n; std::vector<std::vector<float> > tw; };   /* per radix-4 pass: [3][ns] cos, then [3][ns] sin */

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
        int logn = 0; while ((1 << logn) < n) logn++;
        for (int ns = (logn & 1) ? 2 : 1; ns < n; ns *= 4) {
            std::vector<float> t(6 * (size_t)ns);
            for (int k = 0; k < ns; k++)
                for (int r = 1; r < 4; r++) {
                    const double a = -2.0 * M_PI * (double)(r * k) / (double)(4 * ns);
                    t[(size_t)(r - 1) * ns + k] = (float)cos(a); t[(size_t)(3 + r - 1) * ns + k] = (float)sin(a);
                }
            p->tw.push_back(t);
        }
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

/* fp32 Stockham autosort, radix-4 passes (+ one radix-2 pass when log2 n is odd), on split real / imaginary work arrays with
 * per-pass contiguous twiddles so that the inner loops vectorise (8-12 Gflop/s per core with AVX2: the class of speed FFTW
 * reaches, which is what the CPU baseline of bench.py should stand for). */
void fft32(const plan32& p, bool fwd, const cf* in, cf* out)
{
    const int n = p.n;
    if (n == 1) { out[0] = in[0]; return; }
    static thread_local std::vector<float> scratch;
    if (scratch.size() < 4 * (size_t)n) scratch.resize(4 * (size_t)n);
    float* xr = scratch.data(); float* xi = xr + n; float* yr = xi + n; float* yi = yr + n;
    const float sgn = fwd ? 1.0f : -1.0f;       /* conj twiddles + swap +-j for backward */
    const float* fin = reinterpret_cast<const float*>(in);
    for (int i = 0; i < n; i++) { xr[i] = fin[2 * i]; xi[i] = fin[2 * i + 1]; }
    int logn = 0; while ((1 << logn) < n) logn++;
    int ns = 1;
    if (logn & 1) {                              /* radix-2 first (ns = 1, trivial twiddles) */
        const int h = n / 2;
        for (int j = 0; j < h; j++) {
            const float ar = xr[j], ai = xi[j], br = xr[j + h], bi = xi[j + h];
            yr[2 * j] = ar + br; yi[2 * j] = ai + bi; yr[2 * j + 1] = ar - br; yi[2 * j + 1] = ai - bi;
        }
        std::swap(xr, yr); std::swap(xi, yi); ns = 2;
    }
    for (int pass = 0; ns < n; ns *= 4, pass++) {
        const int q = n / 4;
        const float* twr = p.tw[(size_t)pass].data(); const float* twi = twr + 3 * ns;
        for (int j0 = 0; j0 < q; j0 += ns) {
            const float* __restrict__ ar_ = xr + j0;         const float* __restrict__ ai_ = xi + j0;
            const float* __restrict__ br_ = xr + j0 + q;     const float* __restrict__ bi_ = xi + j0 + q;
            const float* __restrict__ cr_ = xr + j0 + 2 * q; const float* __restrict__ ci_ = xi + j0 + 2 * q;
            const float* __restrict__ dr_ = xr + j0 + 3 * q; const float* __restrict__ di_ = xi + j0 + 3 * q;
            float* __restrict__ y0r = yr + 4 * j0; float* __restrict__ y0i = yi + 4 * j0;
            float* __restrict__ y1r = y0r + ns;    float* __restrict__ y1i = y0i + ns;
            float* __restrict__ y2r = y1r + ns;    float* __restrict__ y2i = y1i + ns;
            float* __restrict__ y3r = y2r + ns;    float* __restrict__ y3i = y2i + ns;
#pragma GCC ivdep
            for (int k = 0; k < ns; k++) {
                const float w1r = twr[k], w1i = sgn * twi[k], w2r = twr[ns + k], w2i = sgn * twi[ns + k], w3r = twr[2 * ns + k], w3i = sgn * twi[2 * ns + k];
                const float ar = ar_[k], ai = ai_[k], b0r = br_[k], b0i = bi_[k], c0r = cr_[k], c0i = ci_[k], d0r = dr_[k], d0i = di_[k];
                const float br = b0r * w1r - b0i * w1i, bi = b0r * w1i + b0i * w1r;
                const float cr = c0r * w2r - c0i * w2i, ci = c0r * w2i + c0i * w2r;
                const float dr = d0r * w3r - d0i * w3i, di = d0r * w3i + d0i * w3r;
                const float s0r = ar + cr, s0i = ai + ci, s1r = ar - cr, s1i = ai - ci, s2r = br + dr, s2i = bi + di, s3r = br - dr, s3i = bi - di;
                const float jr = sgn * s3i, ji = -sgn * s3r;                      /* -j*s3 (fwd) / +j*s3 (bwd) */
                y0r[k] = s0r + s2r; y0i[k] = s0i + s2i; y1r[k] = s1r + jr; y1i[k] = s1i + ji;
                y2r[k] = s0r - s2r; y2i[k] = s0i - s2i; y3r[k] = s1r - jr; y3i[k] = s1i - ji;
            }
        }
        std::swap(xr, yr); std::swap(xi, yi);
    }
    float* fo = reinterpret_cast<float*>(out);
    for (int i = 0; i < n; i++) { fo[2 * i] = xr[i]; fo[2 * i + 1] = xi[i]; }
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
