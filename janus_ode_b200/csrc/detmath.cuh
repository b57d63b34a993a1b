// Deterministic fp64 elementary functions (log, exp, pow, tanh) built only from IEEE-754
// basic operations (+ - * / fma, exact bit manipulation).  The same operation sequence gives
// bit-identical results on sm_100a and on any IEEE host, which is what makes bit-level
// parity between the CUDA integrator and its CPU mirror oracle possible: the step-size
// controller of the reference raises `err` and `FacConv` to fractional powers
// (/root/reference/src/cpp/radaute_impl.hpp:569,817,1040-1047,1095; `bpow` of the external
// janus-tensor-dual library) and accepted-step counts are sensitive to the last bit of h.
// Accuracy: log <= ~1 ulp, exp <= ~1 ulp on |t| < 700, pow relative error <= (|y ln x| + 2) eps.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define JB_HD __host__ __device__ __forceinline__
#define JB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define JB_HD inline
#define JB_HD_NOINLINE inline
#endif

namespace jb {

JB_HD int64_t dbl_bits(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  int64_t b;
  std::memcpy(&b, &x, sizeof b);
  return b;
#endif
}

JB_HD double bits_dbl(int64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(b);
#else
  double x;
  std::memcpy(&x, &b, sizeof x);
  return x;
#endif
}

// natural log of a positive normal double
JB_HD double det_log(double x) {
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  int64_t b = dbl_bits(x);
  int e = (int)((b >> 52) & 0x7ff) - 1023;
  double m = bits_dbl((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);  // [1,2)
  if (m > 1.4142135623730951) {
    m = m * 0.5;
    e += 1;
  }
  const double r = (m - 1.0) / (m + 1.0);  // |r| <= 0.1716
  const double r2 = r * r;
  double p = 1.0 / 23.0;
  p = fma(p, r2, 1.0 / 21.0);
  p = fma(p, r2, 1.0 / 19.0);
  p = fma(p, r2, 1.0 / 17.0);
  p = fma(p, r2, 1.0 / 15.0);
  p = fma(p, r2, 1.0 / 13.0);
  p = fma(p, r2, 1.0 / 11.0);
  p = fma(p, r2, 1.0 / 9.0);
  p = fma(p, r2, 1.0 / 7.0);
  p = fma(p, r2, 1.0 / 5.0);
  p = fma(p, r2, 1.0 / 3.0);
  const double t = r + r;
  const double logm = fma(t * r2, p, t);
  const double ed = (double)e;
  return fma(ed, LN2_HI, fma(ed, LN2_LO, logm));
}

// exp(t) for |t| < 700
JB_HD double det_exp(double t) {
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  const double INV_LN2 = 1.44269504088896338700e+00;
  const double kd = floor(fma(t, INV_LN2, 0.5));
  double r = fma(-kd, LN2_HI, t);
  r = fma(-kd, LN2_LO, r);  // |r| <= 0.3466
  double q = 1.0 / 6227020800.0;            // 1/13!
  q = fma(q, r, 1.0 / 479001600.0);         // 1/12!
  q = fma(q, r, 1.0 / 39916800.0);          // 1/11!
  q = fma(q, r, 1.0 / 3628800.0);           // 1/10!
  q = fma(q, r, 1.0 / 362880.0);            // 1/9!
  q = fma(q, r, 1.0 / 40320.0);             // 1/8!
  q = fma(q, r, 1.0 / 5040.0);              // 1/7!
  q = fma(q, r, 1.0 / 720.0);               // 1/6!
  q = fma(q, r, 1.0 / 120.0);               // 1/5!
  q = fma(q, r, 1.0 / 24.0);                // 1/4!
  q = fma(q, r, 1.0 / 6.0);                 // 1/3!
  q = fma(q, r, 0.5);                       // 1/2!
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const int64_t k = (int64_t)kd;
  const double scale = bits_dbl((k + 1023) << 52);  // 2^k, -1022 <= k <= 1023
  return q * scale;
}

// x^y for x > 0.  Not inlined on the device: 5 call sites per attempted step, ~70 instructions each
// -- one shared copy keeps the integrator's code closer to the instruction-cache size.
#ifndef JB_POW_INLINE
JB_HD_NOINLINE double det_pow(double x, double y) { return det_exp(y * det_log(x)); }
#else
JB_HD double det_pow(double x, double y) { return det_exp(y * det_log(x)); }
#endif

// 2^(j/32), j = 0..31, correctly rounded
#define JB_EXP2_TAB_VALUES \
  0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, \
  0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, \
  0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, \
  0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0, \
  0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, \
  0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, \
  0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, \
  0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0 \

static const double DET_EXP2_TAB_H[32] = {JB_EXP2_TAB_VALUES};
#if defined(__CUDACC__)
static __device__ const double DET_EXP2_TAB_D[32] = {JB_EXP2_TAB_VALUES};
#endif
JB_HD const double* det_exp2_tab() {
#if defined(__CUDA_ARCH__)
  return DET_EXP2_TAB_D;
#else
  return DET_EXP2_TAB_H;
#endif
}

// tanh(x) = sign(x) (1 - e) / (1 + e), e = exp(-2|x|), |x| clamped to 20 (beyond it e < 2^-57 and the quotient rounds to
// 1).  Branch-free and short -- the MLP gradient kernel spends three sweeps of 8192 tanh per 64-row tile on it (the first
// version: Horner of degree 13 + floor + IEEE division = 137 SASS instructions per value, 38 % of that kernel):
//   e:  -2|x| = n ln2/32 + r with n = 32 k + j by the 1.5 * 2^52 shifter (no floor, no conversion), |r| <= ln2/64,
//       e = 2^k * tab[j] * P6(r);  2^k goes straight into the exponent field;
//   1 / (1 + e), 1 + e in [1, 2]: quadratic seed (1 % off) + three Newton steps (fma only), then one residual correction of
//       the quotient.  ~26 FP64 operations, ~40 instructions, |error| <= ~2e-16 absolute; the same sequence in
//       oracle/c/radau_oracle.c.  `tab` = det_exp2_tab() or a shared-memory copy of it.  NaN in, NaN out.
JB_HD double det_tanh_t(double x, const double* tab) {
  const double LN2_HI = 6.93147180369123816490e-01;   // 32 significant bits: n * (LN2_HI / 32) is exact
  const double LN2_LO = 1.90821492927058770002e-10;
  const double SHIFT = 6755399441055744.0;            // 1.5 * 2^52
  double ax = fabs(x);
  ax = ax > 20.0 ? 20.0 : ax;
  const double ns = fma(ax, -9.2332482616893657e+01, SHIFT);   // -64 / ln2
  const double nd = ns - SHIFT;
  double r = fma(nd, -(LN2_HI / 32.0), -2.0 * ax);
  r = fma(nd, -(LN2_LO / 32.0), r);
  const int n = (int)(uint32_t)(dbl_bits(ns) & 0xffffffffLL);
  const int j = n & 31, k = n >> 5;
  double p = 1.0 / 720.0;
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double q = p * tab[j];
  const double e = bits_dbl(dbl_bits(q) + ((int64_t)k << 52));
  const double num = 1.0 - e, d = 1.0 + e;
  double y = fma(fma(32.0 / 99.0, d, -48.0 / 33.0), d, 70.0 / 33.0);
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  double th = num * y;
  th = fma(fma(-d, th, num), y, th);
  th = x < 0 ? -th : th;
  return x != x ? x : th;
}
JB_HD double det_tanh(double x) { return det_tanh_t(x, det_exp2_tab()); }

}  // namespace jb
