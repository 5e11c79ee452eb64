// kge_math.h -- scalar math shared by every kernel of libkge_b200.
//
// Everything here is built from IEEE-754 correctly rounded operations only (add, sub, mul, div, sqrt,
// fma, rint), so a CPU restatement with the same constants gives bit-identical results.  That is what
// lets the evaluation path be compared bit-for-bit with the CPU oracle at full size.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KGE_HD __host__ __device__ __forceinline__
#else
#define KGE_HD static inline
#endif

namespace kge {

// ---- un-contractable fp32 primitives (nvcc would otherwise fuse a*b+c) ---------------------------
#if defined(__CUDA_ARCH__)
KGE_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
KGE_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
KGE_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
KGE_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
KGE_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
KGE_HD float fsqrt(float a) { return __fsqrt_rn(a); }
#else
KGE_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
KGE_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
KGE_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
KGE_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
KGE_HD float fdiv(float a, float b) { volatile float r = a / b; return r; }
KGE_HD float fsqrt(float a) { return sqrtf(a); }
#endif

// ---- reproducible sin/cos --------------------------------------------------------------------------
// Cody-Waite reduction by pi/2 (three fp32 constants, exact products for |k| < 2^15; a two-constant
// fp64 reduction beyond that), then the classic degree-7 / degree-8 minimax polynomials on
// [-pi/4, pi/4].  Max observed error vs. libm over [-64 pi, 64 pi]: < 1.5 ulp.
KGE_HD void sincos_rep(float x, float *sn, float *cs) {
  const float TWO_OVER_PI = 0.636619772367581343f;
  float r;
  int q;
  if (fabsf(x) < 40000.0f) {
    float k = rintf(fmul(x, TWO_OVER_PI));
    q = (int)k;
    r = ffma(-k, 1.5703125f, x);                       // pi/2 split: 8 + 11 + 24 significant bits
    r = ffma(-k, 4.837512969970703125e-4f, r);
    r = ffma(-k, 7.54978995489188e-8f, r);
  } else if (fabsf(x) < 1.0e15f) {
    double k = rint((double)x * 0.63661977236758134308);
    q = (int)((long long)k & 3);
    double rd = fma(-k, 1.57079632679489655800e+00, (double)x);
    rd = fma(-k, 6.12323399573676603587e-17, rd);
    r = (float)rd;
  } else {            // inf / nan / astronomically large: follow libm's nan-for-inf convention
    r = x - x;
    q = 0;
  }
  float z = fmul(r, r);
  // sin(r) = r + r z (S1 + z (S2 + z S3))
  float ps = ffma(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = ffma(z, ps, -1.6666654611e-1f);
  float s = ffma(fmul(r, z), ps, r);
  // cos(r) = 1 - z/2 + z^2 (C1 + z (C2 + z C3))
  float pc = ffma(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = ffma(z, pc, 4.166664568298827e-2f);
  float c = ffma(fmul(z, z), pc, ffma(z, -0.5f, 1.0f));
  switch (q & 3) {
    case 0: *sn = s;  *cs = c;  break;
    case 1: *sn = c;  *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}

KGE_HD float sin_rep(float x) {
  float s, c;
  sincos_rep(x, &s, &c);
  return s;
}

}  // namespace kge
