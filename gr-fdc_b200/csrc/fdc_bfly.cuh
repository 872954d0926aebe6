/* fdc_bfly.cuh -- register-resident radix-2/4/8/16 DFT butterflies, complex fp32.
 *
 * DIR = +1 : forward kernel  exp(-j 2 pi t u / R)  (FFTW_FORWARD,  what gr::fft::fft_complex(n, true) runs;
 *            reference call site python/FrequencyDomainChannelizer.py:206)
 * DIR = -1 : backward kernel exp(+j 2 pi t u / R)  (FFTW_BACKWARD, unnormalised; call sites
 *            python/FrequencyDomainChannelizer.py:228, lib/PowerActivationChannel_impl.cc:264-273,
 *            lib/SegmentDetection_impl.cc:404-416)
 * All butterflies work in place on x[0..R) and leave the result in natural order. */
#ifndef FDC_BFLY_CUH
#define FDC_BFLY_CUH
#include "fdc_hd.h"

namespace fdc {

#define FDC_SQRT1_2 0.70710678118654752440f
#define FDC_COS_PI_8 0.92387953251128675613f
#define FDC_SIN_PI_8 0.38268343236508977173f

/* a * W4^1 : forward -j, backward +j */
template <int DIR> FDC_HD float2 rot4(float2 a) { return DIR > 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }
/* a * (c -/+ j s): a * c + rot4(a) * s */
template <int DIR> FDC_HD float2 mulc(float2 a, float c, float s) { return caxpy(rot4<DIR>(a), s, cscale(a, c)); }
/* a * W8^1 = (a + rot4(a)) / sqrt 2,  a * W8^3 = (rot4(a) - a) / sqrt 2 */
template <int DIR> FDC_HD float2 rot8_1(float2 a) { return cscale(cadd(a, rot4<DIR>(a)), FDC_SQRT1_2); }
template <int DIR> FDC_HD float2 rot8_3(float2 a) { return cscale(csub(rot4<DIR>(a), a), FDC_SQRT1_2); }

template <int DIR> FDC_HD void dft2(float2& x0, float2& x1)
{
    const float2 a = x0; x0 = cadd(a, x1); x1 = csub(a, x1);
}
template <int DIR> FDC_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3)
{
    const float2 a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3), a3 = rot4<DIR>(csub(x1, x3));
    x0 = cadd(a0, a2); x1 = cadd(a1, a3); x2 = csub(a0, a2); x3 = csub(a1, a3);
}
template <int DIR> FDC_HD void dft8(float2* x)
{
    /* even / odd quartets, then the radix-2 recombination y[b] = e[b] + W8^b o[b], y[b+4] = e[b] - W8^b o[b] */
    dft4<DIR>(x[0], x[2], x[4], x[6]);
    dft4<DIR>(x[1], x[3], x[5], x[7]);
    const float2 o0 = x[1], o1 = rot8_1<DIR>(x[3]), o2 = rot4<DIR>(x[5]), o3 = rot8_3<DIR>(x[7]);
    const float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, o1); x[5] = csub(e1, o1);
    x[2] = cadd(e2, o2); x[6] = csub(e2, o2);
    x[3] = cadd(e3, o3); x[7] = csub(e3, o3);
}
template <int DIR> FDC_HD void dft16(float2* x)
{
    /* t = a + 4 t1, u = b + 4 c :  X[b+4c] = sum_a W4^{ac} ( W16^{ab} sum_t1 x[a+4t1] W4^{t1 b} ) */
    dft4<DIR>(x[0], x[4], x[8], x[12]);     /* a = 0 : z[0][b] in x[0], x[4], x[8], x[12] */
    dft4<DIR>(x[1], x[5], x[9], x[13]);     /* a = 1 : z[1][b] in x[1+4b] */
    dft4<DIR>(x[2], x[6], x[10], x[14]);
    dft4<DIR>(x[3], x[7], x[11], x[15]);
    /* z[a][b] lives in x[a + 4b]; multiply by W16^{ab} */
    x[5] = mulc<DIR>(x[5], FDC_COS_PI_8, FDC_SIN_PI_8);          /* a1 b1 : W16^1 */
    x[9] = rot8_1<DIR>(x[9]);                                    /* a1 b2 : W16^2 */
    x[13] = mulc<DIR>(x[13], FDC_SIN_PI_8, FDC_COS_PI_8);        /* a1 b3 : W16^3 */
    x[6] = rot8_1<DIR>(x[6]);                                    /* a2 b1 : W16^2 */
    x[10] = rot4<DIR>(x[10]);                                    /* a2 b2 : W16^4 */
    x[14] = rot8_3<DIR>(x[14]);                                  /* a2 b3 : W16^6 */
    x[7] = mulc<DIR>(x[7], FDC_SIN_PI_8, FDC_COS_PI_8);          /* a3 b1 : W16^3 */
    x[11] = rot8_3<DIR>(x[11]);                                  /* a3 b2 : W16^6 */
    x[15] = mulc<DIR>(x[15], -FDC_COS_PI_8, -FDC_SIN_PI_8);      /* a3 b3 : W16^9 = -W16^1 */
    /* for each b: DFT4 over a of x[a + 4b] -> output index c, result X[b + 4c] */
    dft4<DIR>(x[0], x[1], x[2], x[3]);       /* b = 0 : X[0], X[4], X[8], X[12] now in x[0..3] */
    dft4<DIR>(x[4], x[5], x[6], x[7]);       /* b = 1 : X[1], X[5], X[9], X[13] in x[4..7] */
    dft4<DIR>(x[8], x[9], x[10], x[11]);
    dft4<DIR>(x[12], x[13], x[14], x[15]);
    /* x[4b + c] holds X[b + 4c] : transpose the 4x4 register tile into natural order */
#define FDC_SWAP(i, j) { const float2 t_ = x[i]; x[i] = x[j]; x[j] = t_; }
    FDC_SWAP(1, 4) FDC_SWAP(2, 8) FDC_SWAP(3, 12) FDC_SWAP(6, 9) FDC_SWAP(7, 13) FDC_SWAP(11, 14)
#undef FDC_SWAP
}

/* radix 32 = two radix-16 transforms of the even / odd inputs, recombined with W32^k (decimation in time) */
template <int DIR> FDC_HD void dft32(float2* x)
{
    float2 e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
    dft16<DIR>(e); dft16<DIR>(o);
    /* cos / sin of 2 pi k / 32, k = 1..7 */
    const float c[8] = { 1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f,
                         0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f };
    const float sn[8] = { 0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f, 0.70710678118654752440f,
                          0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f };
#pragma unroll
    for (int k = 0; k < 16; k++) {
        float2 t;
        if (k == 0) t = o[0];
        else if (k == 8) t = rot4<DIR>(o[8]);
        else if (k < 8) t = mulc<DIR>(o[k], c[k], sn[k]);                  /* W32^k = c -/+ j s */
        else t = mulc<DIR>(o[k], -sn[k - 8], c[k - 8]);                      /* W32^k, k > 8: cos = -sin(k-8), sin = cos(k-8) */
        x[k] = cadd(e[k], t); x[k + 16] = csub(e[k], t);
    }
}

template <int R, int DIR> struct Bfly;
template <int DIR> struct Bfly<2, DIR> { static FDC_HD void run(float2* x) { dft2<DIR>(x[0], x[1]); } };
template <int DIR> struct Bfly<4, DIR> { static FDC_HD void run(float2* x) { dft4<DIR>(x[0], x[1], x[2], x[3]); } };
template <int DIR> struct Bfly<8, DIR> { static FDC_HD void run(float2* x) { dft8<DIR>(x); } };
template <int DIR> struct Bfly<16, DIR> { static FDC_HD void run(float2* x) { dft16<DIR>(x); } };
template <int DIR> struct Bfly<32, DIR> { static FDC_HD void run(float2* x) { dft32<DIR>(x); } };

}  // namespace fdc
#endif
