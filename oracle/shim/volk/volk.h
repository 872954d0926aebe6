/* Oracle shim (TEST INFRASTRUCTURE ONLY): VOLK's *generic* (non-SIMD) kernels restated.
 * VOLK itself is third-party and absent (version unpinned by the reference).  The loops
 * below follow the published generic implementations: element-wise, strictly sequential
 * accumulation.  SIMD protokernels may sum in a different order; parity tests therefore
 * exclude threshold decisions whose power ratio lies within 1e-5 relative of the threshold. */
#ifndef FDC_SHIM_VOLK_H
#define FDC_SHIM_VOLK_H
#include <complex>
#include <cstdlib>
typedef std::complex<float> lv_32fc_t;
static inline unsigned int volk_get_alignment(void) { return 64; }
static inline void* volk_malloc(size_t size, size_t alignment)
{ void* p = 0; if (posix_memalign(&p, alignment, size ? size : alignment)) return 0; return p; }
static inline void volk_free(void* p) { free(p); }
/* c[i] = a[i]*b[i] = (ar*br - ai*bi, ar*bi + ai*br) in fp32, no FMA contraction (compile with -ffp-contract=off) */
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* c, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n)
{
    for (unsigned int i = 0; i < n; i++) {
        const float ar = a[i].real(), ai = a[i].imag(), br = b[i].real(), bi = b[i].imag();
        c[i] = lv_32fc_t(ar * br - ai * bi, ar * bi + ai * br);
    }
}
static inline void volk_32f_s32f_multiply_32f(float* c, const float* a, const float s, unsigned int n)
{ for (unsigned int i = 0; i < n; i++) c[i] = a[i] * s; }
static inline void volk_32fc_magnitude_squared_32f(float* m, const lv_32fc_t* a, unsigned int n)
{ for (unsigned int i = 0; i < n; i++) { const float r = a[i].real(), q = a[i].imag(); m[i] = r * r + q * q; } }
static inline void volk_32f_accumulator_s32f(float* result, const float* in, unsigned int n)
{ float acc = 0.0f; for (unsigned int i = 0; i < n; i++) acc += in[i]; *result = acc; }
static inline void volk_32f_x2_divide_32f(float* c, const float* a, const float* b, unsigned int n)
{ for (unsigned int i = 0; i < n; i++) c[i] = a[i] / b[i]; }
#endif
