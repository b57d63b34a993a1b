// Third-order jets: value, gradient, Hessian and the symmetric tensor of third derivatives of a scalar H(y).
//
// Replaces the reference's third-order autograd ladders
//   /root/reference/src/cpp/janus_ode_common.hpp:787-2072  (ppppppH, pxppppH, pppxppH, pxpxppH, pppppxH, pxpppxH,
//   pppxpxH, pxpxpxH and their Hu twins: N + N^2 backward passes each)
// which exist to feed evalDynsDual / evalJacDual (:2075-2158, :2302-2416): the sensitivity of the dynamics and of the
// Jacobian to dual directions.  ONE evaluation of H on D3 numbers yields all of them.  Dense (no sparsity masks): this
// is not on the Radau hot path -- the integrator's own sensitivities avoid third derivatives -- it
// serves the stage-evaluation entry points jb200_eval_jac_dual / HamiltonianSystem.jacobian_dual.
//   a*b : v, g_i = a.g_i b.v + a.v b.g_i, h_ij = a.v b.h_ij + a.g_i b.g_j + a.g_j b.g_i + a.h_ij b.v,
//         t_ijk = a.v b.t_ijk + sum_3 a.g b.h + sum_3 a.h b.g + a.t_ijk b.v
//   f(a): g = f1 a.g, h = f1 a.h + f2 a.g a.g, t = f1 a.t + f2 sum_3 a.g a.h + f3 a.g a.g a.g
#pragma once
#include "dual.cuh"

namespace jb {

JB_HD constexpr int tidx(int NV, int i, int j, int k) {  // i <= j <= k
  int idx = 0;
  for (int a = 0; a < i; a++) idx += (NV - a) * (NV - a + 1) / 2;
  for (int b = i; b < j; b++) idx += NV - b;
  return idx + (k - j);
}
JB_HD constexpr int tsorted(int NV, int i, int j, int k) {
  int a = i, b = j, c = k, s = 0;
  if (a > b) { s = a; a = b; b = s; }
  if (b > c) { s = b; b = c; c = s; }
  if (a > b) { s = a; a = b; b = s; }
  return tidx(NV, a, b, c);
}
JB_HD constexpr int hsorted(int NV, int i, int j) { return i <= j ? hidx(NV, i, j) : hidx(NV, j, i); }

template <int NV>
struct D3 {
  static constexpr int nv = NV;
  static constexpr int order = 3;
  static constexpr int NH = NV * (NV + 1) / 2, NT = NV * (NV + 1) * (NV + 2) / 6;
  double v, g[NV], h[NH], t[NT];
};

template <int NV, int I>
JB_HD D3<NV> seed3(double v) {
  D3<NV> r;
  r.v = v;
  for (int i = 0; i < NV; i++) r.g[i] = (i == I) ? 1.0 : 0.0;
  for (int i = 0; i < D3<NV>::NH; i++) r.h[i] = 0.0;
  for (int i = 0; i < D3<NV>::NT; i++) r.t[i] = 0.0;
  return r;
}
template <int NV>
JB_HD D3<NV> const3(double c) {
  D3<NV> r;
  r.v = c;
  for (int i = 0; i < NV; i++) r.g[i] = 0.0;
  for (int i = 0; i < D3<NV>::NH; i++) r.h[i] = 0.0;
  for (int i = 0; i < D3<NV>::NT; i++) r.t[i] = 0.0;
  return r;
}

template <int NV>
JB_HD D3<NV> operator+(const D3<NV>& a, const D3<NV>& b) {
  D3<NV> r;
  r.v = a.v + b.v;
  for (int i = 0; i < NV; i++) r.g[i] = a.g[i] + b.g[i];
  for (int i = 0; i < D3<NV>::NH; i++) r.h[i] = a.h[i] + b.h[i];
  for (int i = 0; i < D3<NV>::NT; i++) r.t[i] = a.t[i] + b.t[i];
  return r;
}
template <int NV>
JB_HD D3<NV> operator-(const D3<NV>& a, const D3<NV>& b) {
  D3<NV> r;
  r.v = a.v - b.v;
  for (int i = 0; i < NV; i++) r.g[i] = a.g[i] - b.g[i];
  for (int i = 0; i < D3<NV>::NH; i++) r.h[i] = a.h[i] - b.h[i];
  for (int i = 0; i < D3<NV>::NT; i++) r.t[i] = a.t[i] - b.t[i];
  return r;
}
template <int NV>
JB_HD D3<NV> operator*(const D3<NV>& a, double c) {
  D3<NV> r;
  r.v = a.v * c;
  for (int i = 0; i < NV; i++) r.g[i] = a.g[i] * c;
  for (int i = 0; i < D3<NV>::NH; i++) r.h[i] = a.h[i] * c;
  for (int i = 0; i < D3<NV>::NT; i++) r.t[i] = a.t[i] * c;
  return r;
}
template <int NV>
JB_HD D3<NV> operator*(double c, const D3<NV>& a) { return a * c; }
template <int NV>
JB_HD D3<NV> operator/(const D3<NV>& a, double c) { return a * (1.0 / c); }
template <int NV>
JB_HD D3<NV> operator-(const D3<NV>& a) { return a * -1.0; }
template <int NV>
JB_HD D3<NV> operator+(const D3<NV>& a, double c) { D3<NV> r = a; r.v = a.v + c; return r; }
template <int NV>
JB_HD D3<NV> operator+(double c, const D3<NV>& a) { return a + c; }
template <int NV>
JB_HD D3<NV> operator-(const D3<NV>& a, double c) { D3<NV> r = a; r.v = a.v - c; return r; }
template <int NV>
JB_HD D3<NV> operator-(double c, const D3<NV>& a) { D3<NV> r = -a; r.v = c - a.v; return r; }

template <int NV>
JB_HD D3<NV> operator*(const D3<NV>& a, const D3<NV>& b) {
  D3<NV> r;
  r.v = a.v * b.v;
  for (int i = 0; i < NV; i++) r.g[i] = fma(a.g[i], b.v, a.v * b.g[i]);
  for (int i = 0; i < NV; i++)
    for (int j = i; j < NV; j++) {
      const int k = hidx(NV, i, j);
      r.h[k] = fma(a.h[k], b.v, fma(a.g[j], b.g[i], fma(a.g[i], b.g[j], a.v * b.h[k])));
    }
  for (int i = 0; i < NV; i++)
    for (int j = i; j < NV; j++)
      for (int k = j; k < NV; k++) {
        const int n = tidx(NV, i, j, k), hij = hidx(NV, i, j), hik = hidx(NV, i, k), hjk = hidx(NV, j, k);
        double s = a.v * b.t[n];
        s = fma(a.g[i], b.h[hjk], s);
        s = fma(a.g[j], b.h[hik], s);
        s = fma(a.g[k], b.h[hij], s);
        s = fma(a.h[hij], b.g[k], s);
        s = fma(a.h[hik], b.g[j], s);
        s = fma(a.h[hjk], b.g[i], s);
        r.t[n] = fma(a.t[n], b.v, s);
      }
  return r;
}

template <int NV>
JB_HD D3<NV> chain3(const D3<NV>& a, double f0, double f1, double f2, double f3) {
  D3<NV> r;
  r.v = f0;
  for (int i = 0; i < NV; i++) r.g[i] = f1 * a.g[i];
  for (int i = 0; i < NV; i++)
    for (int j = i; j < NV; j++) {
      const int k = hidx(NV, i, j);
      r.h[k] = fma(f2, a.g[i] * a.g[j], f1 * a.h[k]);
    }
  for (int i = 0; i < NV; i++)
    for (int j = i; j < NV; j++)
      for (int k = j; k < NV; k++) {
        const int n = tidx(NV, i, j, k);
        const double m = fma(a.g[i], a.h[hidx(NV, j, k)], fma(a.g[j], a.h[hidx(NV, i, k)], a.g[k] * a.h[hidx(NV, i, j)]));
        r.t[n] = fma(f3, a.g[i] * a.g[j] * a.g[k], fma(f2, m, f1 * a.t[n]));
      }
  return r;
}

// functions (more specialised than the generic D1 / D2 versions of dual.cuh, so these win for D3)
template <int NV>
JB_HD D3<NV> recip(const D3<NV>& a) {
  const double f = 1.0 / a.v, f1 = -(f * f), f2 = -2.0 * (f1 * f), f3 = -3.0 * (f2 * f);
  return chain3(a, f, f1, f2, f3);
}
template <int NV>
JB_HD D3<NV> operator/(const D3<NV>& a, const D3<NV>& b) { return a * recip(b); }
template <int NV>
JB_HD D3<NV> operator/(double c, const D3<NV>& b) { return recip(b) * c; }
template <int NV>
JB_HD D3<NV> square(const D3<NV>& a) { return a * a; }
template <int NV>
JB_HD D3<NV> sqrt(const D3<NV>& a) {
  const double f = ::sqrt(a.v), f1 = 0.5 / f, f2 = -0.5 * (f1 / a.v), f3 = -1.5 * (f2 / a.v);
  return chain3(a, f, f1, f2, f3);
}
template <int NV>
JB_HD D3<NV> exp(const D3<NV>& a) {
  const double f = det_exp(a.v);
  return chain3(a, f, f, f, f);
}
template <int NV>
JB_HD D3<NV> log(const D3<NV>& a) {
  const double f = det_log(a.v), f1 = 1.0 / a.v, f2 = -(f1 * f1), f3 = -2.0 * (f2 * f1);
  return chain3(a, f, f1, f2, f3);
}
template <int NV>
JB_HD D3<NV> tanh(const D3<NV>& a) {
  const double f = det_tanh(a.v), f1 = 1.0 - f * f, f2 = -2.0 * (f * f1), f3 = -2.0 * (f1 * f1 + f * f2);
  return chain3(a, f, f1, f2, f3);
}
template <int NV>
JB_HD D3<NV> sin(const D3<NV>& a) {
  const double s = ::sin(a.v), c = ::cos(a.v);
  return chain3(a, s, c, -s, -c);
}
template <int NV>
JB_HD D3<NV> cos(const D3<NV>& a) {
  const double s = ::sin(a.v), c = ::cos(a.v);
  return chain3(a, c, -s, -c, s);
}

}  // namespace jb
