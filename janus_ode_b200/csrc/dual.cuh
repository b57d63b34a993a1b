// Compile-time-sparse dual (first order) and hyper-dual (second order) numbers.
//
// Replaces the reference's autograd derivative ladders
//   /root/reference/src/cpp/janus_ode_common.hpp:116-134 (safe_jac), :177-279 (pxH, ppH),
//   :282-718 (ppppH, pxpxH, pxppH, pppxH)
// and the external janus-tensor-dual TensorDual arithmetic for the value+gradient(+Hessian)
// of a scalar Hamiltonian H(y), y = [p, x].  One evaluation of H on D1 numbers gives the
// whole gradient (the dynamics), one evaluation on D2 numbers gives the Hessian (the Jacobian).
//
// Each number carries, as template parameters, the bit mask GM of variables it depends on
// and (D2) the mask HM of non-zero Hessian pairs.  Operations only touch entries whose mask
// bit is set, so after inlining the structurally-zero partials cost nothing -- the work per
// evaluation is that of hand-derived derivatives -- while the arithmetic stays a fixed,
// documented sequence (ARITHMETIC CONTRACT below) that a dense CPU implementation reproduces
// bit for bit (adding an exact zero / multiplying by an exact zero changes nothing).
//
// ARITHMETIC CONTRACT (mirrored by oracle/c/radau_oracle.c, "mirror" mode):
//   a+b, a-b  : componentwise; an entry present on one side only is copied (negated for -b)
//   a*b       : v = a.v*b.v
//               g_i  = fma(a.g_i, b.v, a.v*b.g_i)          [a.g_i*b.v / a.v*b.g_i if one side]
//               h_ij = terms, in this order, first present term a plain product, the others
//                      fma-accumulated:  a.v*b.h_ij , a.g_i*b.g_j , a.g_j*b.g_i , a.h_ij*b.v
//   f(a)      : v = f, g_i = f1*a.g_i, h_ij = f1*a.h_ij then fma(f2, a.g_i*a.g_j, .)
//   a/b       : a * recip(b), recip: f = 1/v, f1 = -(f*f), f2 = -2*(f1*f)
//   a*c, c*a, a/c (c double): every entry times c (a/c multiplies by r = 1.0/c)
#pragma once
#include "detmath.cuh"

namespace jb {

JB_HD constexpr int hidx(int NV, int i, int j) {  // i <= j, row-major upper triangle
  return i * NV - (i * (i - 1)) / 2 + (j - i);
}
JB_HD constexpr unsigned long long hbit(int NV, int i, int j) {
  return 1ull << (i <= j ? hidx(NV, i, j) : hidx(NV, j, i));
}
// pairs (i<=j) with (ga_i & gb_j) | (ga_j & gb_i)
JB_HD constexpr unsigned long long outer_mask(int NV, unsigned GA, unsigned GB) {
  unsigned long long m = 0;
  for (int i = 0; i < NV; i++)
    for (int j = i; j < NV; j++)
      if ((((GA >> i) & 1u) && ((GB >> j) & 1u)) || (((GA >> j) & 1u) && ((GB >> i) & 1u)))
        m |= 1ull << hidx(NV, i, j);
  return m;
}

// ------------------------------------------------------------------ first order
template <int NV, unsigned GM>
struct D1 {
  static constexpr int nv = NV;
  static constexpr unsigned gm = GM;
  static constexpr unsigned long long hm = 0ull;
  static constexpr int order = 1;
  double v;
  double g[NV];
};

template <int NV, int I>
JB_HD D1<NV, (1u << I)> seed1(double v) {
  D1<NV, (1u << I)> r;
  r.v = v;
  r.g[I] = 1.0;
  return r;
}

// ------------------------------------------------------------------ second order
template <int NV, unsigned GM, unsigned long long HM>
struct D2 {
  static constexpr int nv = NV;
  static constexpr unsigned gm = GM;
  static constexpr unsigned long long hm = HM;
  static constexpr int order = 2;
  double v;
  double g[NV];
  double h[NV * (NV + 1) / 2];
};

template <int NV, int I>
JB_HD D2<NV, (1u << I), 0ull> seed2(double v) {
  D2<NV, (1u << I), 0ull> r;
  r.v = v;
  r.g[I] = 1.0;
  return r;
}

#define JB_BIT(m, i) ((((m) >> (i)) & 1) != 0)

// ------------------------------------------------------------------ D1 operators
template <int NV, unsigned GA, unsigned GB>
JB_HD D1<NV, GA | GB> operator+(const D1<NV, GA>& a, const D1<NV, GB>& b) {
  D1<NV, GA | GB> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (JB_BIT(GA, i) && JB_BIT(GB, i)) r.g[i] = a.g[i] + b.g[i];
    else if (JB_BIT(GA, i)) r.g[i] = a.g[i];
    else if (JB_BIT(GB, i)) r.g[i] = b.g[i];
  }
  return r;
}
template <int NV, unsigned GA, unsigned GB>
JB_HD D1<NV, GA | GB> operator-(const D1<NV, GA>& a, const D1<NV, GB>& b) {
  D1<NV, GA | GB> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (JB_BIT(GA, i) && JB_BIT(GB, i)) r.g[i] = a.g[i] - b.g[i];
    else if (JB_BIT(GA, i)) r.g[i] = a.g[i];
    else if (JB_BIT(GB, i)) r.g[i] = -b.g[i];
  }
  return r;
}
template <int NV, unsigned GA, unsigned GB>
JB_HD D1<NV, GA | GB> operator*(const D1<NV, GA>& a, const D1<NV, GB>& b) {
  D1<NV, GA | GB> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (JB_BIT(GA, i) && JB_BIT(GB, i)) r.g[i] = fma(a.g[i], b.v, a.v * b.g[i]);
    else if (JB_BIT(GA, i)) r.g[i] = a.g[i] * b.v;
    else if (JB_BIT(GB, i)) r.g[i] = a.v * b.g[i];
  }
  return r;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator-(const D1<NV, GA>& a) {
  D1<NV, GA> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (JB_BIT(GA, i)) r.g[i] = -a.g[i];
  return r;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator+(const D1<NV, GA>& a, double c) {
  D1<NV, GA> r = a;
  r.v = a.v + c;
  return r;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator+(double c, const D1<NV, GA>& a) {
  D1<NV, GA> r = a;
  r.v = c + a.v;
  return r;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator-(const D1<NV, GA>& a, double c) {
  D1<NV, GA> r = a;
  r.v = a.v - c;
  return r;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator-(double c, const D1<NV, GA>& a) {
  D1<NV, GA> r;
  r.v = c - a.v;
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (JB_BIT(GA, i)) r.g[i] = -a.g[i];
  return r;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator*(const D1<NV, GA>& a, double c) {
  D1<NV, GA> r;
  r.v = a.v * c;
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (JB_BIT(GA, i)) r.g[i] = a.g[i] * c;
  return r;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator*(double c, const D1<NV, GA>& a) {
  return a * c;
}
template <int NV, unsigned GA>
JB_HD D1<NV, GA> operator/(const D1<NV, GA>& a, double c) {
  return a * (1.0 / c);
}
// f(a) with value f0 and derivative f1 at a.v
template <int NV, unsigned GA>
JB_HD D1<NV, GA> chain(const D1<NV, GA>& a, double f0, double f1, double /*f2*/) {
  D1<NV, GA> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (JB_BIT(GA, i)) r.g[i] = f1 * a.g[i];
  return r;
}

// ------------------------------------------------------------------ D2 operators
template <int NV, unsigned GA, unsigned long long HA, unsigned GB, unsigned long long HB>
JB_HD D2<NV, GA | GB, HA | HB> operator+(const D2<NV, GA, HA>& a, const D2<NV, GB, HB>& b) {
  D2<NV, GA | GB, HA | HB> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (JB_BIT(GA, i) && JB_BIT(GB, i)) r.g[i] = a.g[i] + b.g[i];
    else if (JB_BIT(GA, i)) r.g[i] = a.g[i];
    else if (JB_BIT(GB, i)) r.g[i] = b.g[i];
  }
#pragma unroll
  for (int k = 0; k < NV * (NV + 1) / 2; k++) {
    if (JB_BIT(HA, k) && JB_BIT(HB, k)) r.h[k] = a.h[k] + b.h[k];
    else if (JB_BIT(HA, k)) r.h[k] = a.h[k];
    else if (JB_BIT(HB, k)) r.h[k] = b.h[k];
  }
  return r;
}
template <int NV, unsigned GA, unsigned long long HA, unsigned GB, unsigned long long HB>
JB_HD D2<NV, GA | GB, HA | HB> operator-(const D2<NV, GA, HA>& a, const D2<NV, GB, HB>& b) {
  D2<NV, GA | GB, HA | HB> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (JB_BIT(GA, i) && JB_BIT(GB, i)) r.g[i] = a.g[i] - b.g[i];
    else if (JB_BIT(GA, i)) r.g[i] = a.g[i];
    else if (JB_BIT(GB, i)) r.g[i] = -b.g[i];
  }
#pragma unroll
  for (int k = 0; k < NV * (NV + 1) / 2; k++) {
    if (JB_BIT(HA, k) && JB_BIT(HB, k)) r.h[k] = a.h[k] - b.h[k];
    else if (JB_BIT(HA, k)) r.h[k] = a.h[k];
    else if (JB_BIT(HB, k)) r.h[k] = -b.h[k];
  }
  return r;
}
template <int NV, unsigned GA, unsigned long long HA, unsigned GB, unsigned long long HB>
JB_HD D2<NV, GA | GB, HA | HB | outer_mask(NV, GA, GB)> operator*(const D2<NV, GA, HA>& a,
                                                                  const D2<NV, GB, HB>& b) {
  D2<NV, GA | GB, HA | HB | outer_mask(NV, GA, GB)> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (JB_BIT(GA, i) && JB_BIT(GB, i)) r.g[i] = fma(a.g[i], b.v, a.v * b.g[i]);
    else if (JB_BIT(GA, i)) r.g[i] = a.g[i] * b.v;
    else if (JB_BIT(GB, i)) r.g[i] = a.v * b.g[i];
  }
#pragma unroll
  for (int i = 0; i < NV; i++) {
#pragma unroll
    for (int j = i; j < NV; j++) {
      const int k = hidx(NV, i, j);
      bool have = false;
      double t = 0.0;
      if (JB_BIT(HB, k)) { t = a.v * b.h[k]; have = true; }
      if (JB_BIT(GA, i) && JB_BIT(GB, j)) { t = have ? fma(a.g[i], b.g[j], t) : a.g[i] * b.g[j]; have = true; }
      if (JB_BIT(GA, j) && JB_BIT(GB, i)) { t = have ? fma(a.g[j], b.g[i], t) : a.g[j] * b.g[i]; have = true; }
      if (JB_BIT(HA, k)) { t = have ? fma(a.h[k], b.v, t) : a.h[k] * b.v; have = true; }
      if (have) r.h[k] = t;
    }
  }
  return r;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator-(const D2<NV, GA, HA>& a) {
  D2<NV, GA, HA> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (JB_BIT(GA, i)) r.g[i] = -a.g[i];
#pragma unroll
  for (int k = 0; k < NV * (NV + 1) / 2; k++)
    if (JB_BIT(HA, k)) r.h[k] = -a.h[k];
  return r;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator+(const D2<NV, GA, HA>& a, double c) {
  D2<NV, GA, HA> r = a;
  r.v = a.v + c;
  return r;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator+(double c, const D2<NV, GA, HA>& a) {
  D2<NV, GA, HA> r = a;
  r.v = c + a.v;
  return r;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator-(const D2<NV, GA, HA>& a, double c) {
  D2<NV, GA, HA> r = a;
  r.v = a.v - c;
  return r;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator-(double c, const D2<NV, GA, HA>& a) {
  D2<NV, GA, HA> r = -a;
  r.v = c - a.v;
  return r;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator*(const D2<NV, GA, HA>& a, double c) {
  D2<NV, GA, HA> r;
  r.v = a.v * c;
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (JB_BIT(GA, i)) r.g[i] = a.g[i] * c;
#pragma unroll
  for (int k = 0; k < NV * (NV + 1) / 2; k++)
    if (JB_BIT(HA, k)) r.h[k] = a.h[k] * c;
  return r;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator*(double c, const D2<NV, GA, HA>& a) {
  return a * c;
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA> operator/(const D2<NV, GA, HA>& a, double c) {
  return a * (1.0 / c);
}
template <int NV, unsigned GA, unsigned long long HA>
JB_HD D2<NV, GA, HA | outer_mask(NV, GA, GA)> chain(const D2<NV, GA, HA>& a, double f0, double f1,
                                                    double f2) {
  D2<NV, GA, HA | outer_mask(NV, GA, GA)> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (JB_BIT(GA, i)) r.g[i] = f1 * a.g[i];
#pragma unroll
  for (int i = 0; i < NV; i++) {
#pragma unroll
    for (int j = i; j < NV; j++) {
      const int k = hidx(NV, i, j);
      bool have = false;
      double t = 0.0;
      if (JB_BIT(HA, k)) { t = f1 * a.h[k]; have = true; }
      if (JB_BIT(GA, i) && JB_BIT(GA, j)) {
        const double gg = a.g[i] * a.g[j];
        t = have ? fma(f2, gg, t) : f2 * gg;
        have = true;
      }
      if (have) r.h[k] = t;
    }
  }
  return r;
}

// ------------------------------------------------------------------ functions (both orders)
template <class A>
JB_HD auto recip(const A& a) {
  const double f = 1.0 / a.v;
  const double f1 = -(f * f);
  const double f2 = -2.0 * (f1 * f);
  return chain(a, f, f1, f2);
}
template <class A, class B, int = A::order, int = B::order>
JB_HD auto operator/(const A& a, const B& b) {
  return a * recip(b);
}
template <class A, int = A::order>
JB_HD auto operator/(double c, const A& b) {
  return recip(b) * c;
}
template <class A>
JB_HD auto square(const A& a) {
  return a * a;
}
template <class A>
JB_HD auto sqrt(const A& a) {
  const double f = ::sqrt(a.v);
  const double f1 = 0.5 / f;
  const double f2 = -0.5 * (f1 / a.v);
  return chain(a, f, f1, f2);
}
template <class A>
JB_HD auto exp(const A& a) {  // deterministic exp
  const double f = det_exp(a.v);
  return chain(a, f, f, f);
}
template <class A>
JB_HD auto log(const A& a) {  // deterministic log
  const double f = det_log(a.v);
  const double f1 = 1.0 / a.v;
  return chain(a, f, f1, -(f1 * f1));
}
template <class A>
JB_HD auto tanh(const A& a) {  // deterministic tanh
  const double f = det_tanh(a.v);
  const double f1 = 1.0 - f * f;
  return chain(a, f, f1, -2.0 * (f * f1));
}
template <class A>
JB_HD auto sin(const A& a) {  // libm/libdevice sin, cos: tolerance-level parity only
  const double s = ::sin(a.v), c = ::cos(a.v);
  return chain(a, s, c, -s);
}
template <class A>
JB_HD auto cos(const A& a) {
  const double s = ::sin(a.v), c = ::cos(a.v);
  return chain(a, c, -s, -c);
}

// value helpers usable on plain doubles too (control laws act on values)
JB_HD double value(double x) { return x; }
template <class A, int = A::order>
JB_HD double value(const A& a) {
  return a.v;
}

}  // namespace jb
