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

// ---- tables (run-time indexed: global memory on the device, L1-resident; the MLP tile kernel keeps its own copy of the
// exp2 table in shared memory)
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

// log: 1 / c_j and -log(1 / c_j) for c_j = 1 + (j + 1/2) / 32, j = 0..31 (1 / c_j rounded first, its log taken exactly)
#define JB_LOG_INVC_VALUES \
  0x1.f81f81f81f820p-1, 0x1.e9131abf0b767p-1, 0x1.dae6076b981dbp-1, 0x1.cd85689039b0bp-1, \
  0x1.c0e070381c0e0p-1, 0x1.b4e81b4e81b4fp-1, 0x1.a98ef606a63bep-1, 0x1.9ec8e951033d9p-1, \
  0x1.948b0fcd6e9e0p-1, 0x1.8acb90f6bf3aap-1, 0x1.8181818181818p-1, 0x1.78a4c8178a4c8p-1, \
  0x1.702e05c0b8170p-1, 0x1.6816816816817p-1, 0x1.6058160581606p-1, 0x1.58ed2308158edp-1, \
  0x1.51d07eae2f815p-1, 0x1.4afd6a052bf5bp-1, 0x1.446f86562d9fbp-1, 0x1.3e22cbce4a902p-1, \
  0x1.3813813813814p-1, 0x1.323e34a2b10bfp-1, 0x1.2c9fb4d812ca0p-1, 0x1.27350b8812735p-1, \
  0x1.21fb78121fb78p-1, 0x1.1cf06ada2811dp-1, 0x1.1811811811812p-1, 0x1.135c81135c811p-1, \
  0x1.0ecf56be69c90p-1, 0x1.0a6810a6810a7p-1, 0x1.0624dd2f1a9fcp-1, 0x1.0204081020408p-1
#define JB_LOG_LOGC_VALUES \
  0x1.fc0a8b0fc03c4p-7, 0x1.77458f632dcffp-5, 0x1.341d7961bd1d0p-4, 0x1.a926d3a4ad562p-4, \
  0x1.0d77e7cd08e5bp-3, 0x1.44d2b6ccb7d1cp-3, 0x1.7ab890210d907p-3, 0x1.af3c94e80bff3p-3, \
  0x1.e27076e2af2e8p-3, 0x1.0a324e27390e2p-2, 0x1.22941fbcf7966p-2, 0x1.3a64c556945eap-2, \
  0x1.51aad872df82ep-2, 0x1.686c81e9b14adp-2, 0x1.7eaf83b82afc2p-2, 0x1.947941c2116fbp-2, \
  0x1.a9cec9a9a084ap-2, 0x1.beb4d9da71b7ap-2, 0x1.d32fe7e00ebd5p-2, 0x1.e744261d68789p-2, \
  0x1.faf588f78f31dp-2, 0x1.0723e5c1cdf41p-1, 0x1.109f39e2d4c96p-1, 0x1.19ee6b467c96fp-1, \
  0x1.23130d7bebf43p-1, 0x1.2c0e9ed448e8cp-1, 0x1.34e289d9ce1d2p-1, 0x1.3d9026a7156fbp-1, \
  0x1.4618bc21c5ec2p-1, 0x1.4e7d811b75bb0p-1, 0x1.56bf9d5b3f399p-1, 0x1.5ee02a9241676p-1

static const double DET_LOG_INVC_H[32] = {JB_LOG_INVC_VALUES};
static const double DET_LOG_LOGC_H[32] = {JB_LOG_LOGC_VALUES};
#if defined(__CUDACC__)
static __device__ const double DET_LOG_INVC_D[32] = {JB_LOG_INVC_VALUES};
static __device__ const double DET_LOG_LOGC_D[32] = {JB_LOG_LOGC_VALUES};
#endif

// natural log of a positive normal double.  x = 2^e m, m in [1, 2), j = the top five mantissa bits:
//   r = m / c_j - 1 by ONE fma with the tabulated 1 / c_j (|r| <= 1 / 64), log m = log c_j + log1p(r), log1p by a polynomial
//   of degree 8 -- no division, no branch (the first version: (m - 1) / (m + 1) by an IEEE division + 11 Horner terms).
// Absolute error <= ~1.5e-16 max(1, |log x|): what pow needs (its relative error is y times that).
JB_HD double det_log(double x) {
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  const int64_t b = dbl_bits(x);
  const int e = (int)((b >> 52) & 0x7ff) - 1023;
  const int j = (int)((b >> 47) & 31);
  const double m = bits_dbl((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);  // [1,2)
#if defined(__CUDA_ARCH__)
  const double invc = DET_LOG_INVC_D[j], logc = DET_LOG_LOGC_D[j];
#else
  const double invc = DET_LOG_INVC_H[j], logc = DET_LOG_LOGC_H[j];
#endif
  const double r = fma(m, invc, -1.0);
  double p = -1.0 / 8.0;
  p = fma(p, r, 1.0 / 7.0);
  p = fma(p, r, -1.0 / 6.0);
  p = fma(p, r, 1.0 / 5.0);
  p = fma(p, r, -1.0 / 4.0);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  const double ed = (double)e;
  const double hi = fma(ed, LN2_HI, logc);
  const double lo = fma(r * r, p, fma(ed, LN2_LO, r));
  return hi + lo;
}

// exp(t) for |t| < 700: t = n ln2/32 + r with n = 32 k + j by the 1.5 * 2^52 shifter, |r| <= ln2/64,
// exp t = 2^k tab[j] P6(r), 2^k added into the exponent field (the first version: floor + Horner of degree 13)
JB_HD double det_exp(double t) {
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  const double SHIFT = 6755399441055744.0;            // 1.5 * 2^52
  const double ns = fma(t, 4.6166241308446828e+01, SHIFT);   // 32 / ln2
  const double nd = ns - SHIFT;
  double r = fma(nd, -(LN2_HI / 32.0), t);
  r = fma(nd, -(LN2_LO / 32.0), r);
  const int n = (int)(uint32_t)(dbl_bits(ns) & 0xffffffffLL);
  const int j = n & 31, k = n >> 5;
  double q = 1.0 / 720.0;
  q = fma(q, r, 1.0 / 120.0);
  q = fma(q, r, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  q = q * det_exp2_tab()[j];
  return bits_dbl(dbl_bits(q) + ((int64_t)k << 52));
}

// x^y for x > 0.  Not inlined on the device: 5 call sites per attempted step, ~70 instructions each
// -- one shared copy keeps the integrator's code closer to the instruction-cache size.
#ifndef JB_POW_INLINE
JB_HD_NOINLINE double det_pow(double x, double y) { return det_exp(y * det_log(x)); }
#else
JB_HD double det_pow(double x, double y) { return det_exp(y * det_log(x)); }
#endif

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
