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

/* ---- complex helpers ----------------------------------------------------------------------
 * On sm_100a a complex fp32 value is one 64-bit register pair and the packed instructions FADD2 / FMUL2 / FFMA2
 * (PTX add/mul/fma.rn.f32x2) work on both halves at once: a complex add is ONE issue slot instead of two, a complex
 * multiply two instead of four.  Their operands take a half swap (.LO_HI), per-half signs and a scalar broadcast for
 * free, so multiplications by +-j and (c -+ js) need no extra moves: ptxas folds the mov.b64 {..} packing, the
 * negations and the swaps written below into operand modifiers (checked in the SASS, tools/sass_summary.py).
 * The packed ops have the lane throughput of the scalar ones (measured, tools/ubench.cu: 3.8 warp-instructions/ns/SM
 * against 7.5) -- what they save is issue slots, which is what the FFT kernels are short of.
 * Rounding is identical to the scalar forms (every half is an IEEE add / mul / fma, round to nearest). */
#if defined(__CUDA_ARCH__) && !defined(FDC_NO_F32X2)
#define FDC_F32X2 1
FDC_D float2 f2_add(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
FDC_D float2 f2_mul(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
FDC_D float2 f2_fma(float2 a, float2 b, float2 c)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
FDC_HD float2 cadd(float2 a, float2 b) { return f2_add(a, b); }
FDC_HD float2 csub(float2 a, float2 b) { return f2_add(a, make_float2(-b.x, -b.y)); }
/* a * s, s real */
FDC_HD float2 cscale(float2 a, float s) { return f2_mul(a, make_float2(s, s)); }
/* a * (wr + j wi) = a * wr + (-a.y, a.x) * wi */
FDC_HD float2 cmul(float2 a, float2 w) { return f2_fma(make_float2(-a.y, a.x), make_float2(w.y, w.y), f2_mul(a, make_float2(w.x, w.x))); }
/* a * conj(w) = a * wr + (a.y, -a.x) * wi */
FDC_HD float2 cmulc(float2 a, float2 w) { return f2_fma(make_float2(a.y, -a.x), make_float2(w.y, w.y), f2_mul(a, make_float2(w.x, w.x))); }
/* acc + a * s, s real */
FDC_HD float2 caxpy(float2 a, float s, float2 acc) { return f2_fma(a, make_float2(s, s), acc); }
#else
FDC_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FDC_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FDC_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
FDC_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FDC_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
FDC_HD float2 caxpy(float2 a, float s, float2 acc) { return make_float2(a.x * s + acc.x, a.y * s + acc.y); }
#endif
/* VOLK-generic complex multiply: (ar*br - ai*bi, ar*bi + ai*br), four products and two sums
 * each rounded on its own (volk_32fc_x2_multiply_32fc generic kernel; reference call sites
 * lib/phase_shifting_windowing_vcc_impl.cc:81, lib/PowerActivationChannel_impl.cc:267,
 * lib/SegmentDetection_impl.cc:407-410).  Kept in scalar form: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
 * (seen in the SASS and caught by test_psw_block_bit_exact), the scalar __fmul_rn / __fadd_rn are never contracted. */
FDC_HD float2 cmul_exact(float2 a, float2 b)
{
    return make_float2(fdc_sub(fdc_mul(a.x, b.x), fdc_mul(a.y, b.y)), fdc_add(fdc_mul(a.x, b.y), fdc_mul(a.y, b.x)));
}

#endif
