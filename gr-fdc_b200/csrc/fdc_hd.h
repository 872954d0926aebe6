/* fdc_hd.h -- host/device portability layer for the FDC kernels.
 *
 * The kernels are written as "phases": plain inline functions of (thread id, per-thread
 * register array, shared-memory pointer).  nvcc compiles them into the sm_100a kernels in
 * fdc_kernels.cu; g++ can compile the very same functions (FDC_HOST_EMU) so that the CPU
 * test-suite can step every phase over all thread ids and check the index arithmetic
 * before GPU time is spent (tests/emu/).  The emulation build is test infrastructure only,
 * it is never part of libfdc_b200.so. */
#ifndef FDC_HD_H
#define FDC_HD_H

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define FDC_HD __host__ __device__ __forceinline__
#define FDC_D __device__ __forceinline__
#else
#include <cmath>
#define FDC_HD inline
#define FDC_D inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

#include <stdint.h>

/* ---- exactly-rounded fp32 ops (never contracted into FMA), used where the reference's
 *      arithmetic order is mirrored (window multiply, |x|^2 sums, ratios) ---------------- */
#if defined(__CUDA_ARCH__)
FDC_HD float fdc_mul(float a, float b) { return __fmul_rn(a, b); }
FDC_HD float fdc_add(float a, float b) { return __fadd_rn(a, b); }
FDC_HD float fdc_sub(float a, float b) { return __fsub_rn(a, b); }
FDC_HD float fdc_div(float a, float b) { return __fdiv_rn(a, b); }
template <class T> FDC_HD T fdc_ldg(const T* p) { return __ldg(p); }
#else
/* host: translation units are compiled with -ffp-contract=off */
FDC_HD float fdc_mul(float a, float b) { return a * b; }
FDC_HD float fdc_add(float a, float b) { return a + b; }
FDC_HD float fdc_sub(float a, float b) { return a - b; }
FDC_HD float fdc_div(float a, float b) { return a / b; }
template <class T> FDC_HD T fdc_ldg(const T* p) { return *p; }
#endif

/* complex helpers (free to contract: used inside the FFT butterflies) */
FDC_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FDC_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FDC_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
/* VOLK-generic complex multiply: (ar*br - ai*bi, ar*bi + ai*br), four products and two sums
 * each rounded on its own (volk_32fc_x2_multiply_32fc generic kernel; reference call sites
 * lib/phase_shifting_windowing_vcc_impl.cc:81, lib/PowerActivationChannel_impl.cc:267,
 * lib/SegmentDetection_impl.cc:407-410). */
FDC_HD float2 cmul_exact(float2 a, float2 b)
{
    return make_float2(fdc_sub(fdc_mul(a.x, b.x), fdc_mul(a.y, b.y)), fdc_add(fdc_mul(a.x, b.y), fdc_mul(a.y, b.x)));
}

#endif
