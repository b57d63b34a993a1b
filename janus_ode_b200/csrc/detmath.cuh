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

// tanh via exp(-2|x|): (1 - e)/(1 + e), sign restored; exact +-1 saturation for |x| > 20
JB_HD double det_tanh(double x) {
  const double ax = fabs(x);
  if (ax > 20.0) return x > 0 ? 1.0 : -1.0;
  const double e = det_exp(-2.0 * ax);
  const double th = (1.0 - e) / (1.0 + e);
  return x < 0 ? -th : th;
}

}  // namespace jb
