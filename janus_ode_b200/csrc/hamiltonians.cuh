// Built-in Hamiltonians and the H -> (dynamics, Jacobian) adaptor.
//
// Replaces /root/reference/src/cpp/janus_ode_common.hpp:2161-2217 (evalDyns / evalDynsU) and
// :2220-2299 (evalJac / evalJacU).  State layout y = [p(0:N), x(N:2N)] (:2167-2169):
//   ydot = [ sigma * dH/dx ; dH/dp ]      sigma = +1 reference (:2174-2184), -1 canonical
//   J    = d ydot / d y = [[pppxH, pxpxH], [ppppH, pxppH]] with the sigma on the pdot rows.
// A supplied control u = control(y) is evaluated on VALUES and held constant under
// differentiation (evalDynsU/evalJacU semantics, :2188-2217, :2259-2299).
//
// A Hamiltonian is a struct with
//   N, NV = 2N, NPAR, HAS_CONTROL
//   template <class... V, class PAR> static auto H(v0, ..., v_{NV-1}, const PAR& par, double u)
//        arguments in y-order (p_1..p_N, x_1..x_N), each a D1/D2 number with its own mask; par[k] is a plain double
//        (ParPlain) or, for parameter sensitivities, a hyper-dual number (ParDual): take it with `const auto`
//   static double control(const double* y, const double* par)          (if HAS_CONTROL)
//   FLOPS_F, FLOPS_J   FP64 instructions of one dynamics / Jacobian evaluation, counted from the
//                      SASS of dyns_kernel / jac_kernel (roofline accounting only)
// User Hamiltonians compiled at run time (NVRTC) use exactly this interface.
#pragma once
#include <utility>

#include "dual.cuh"
#include "dual3.cuh"

namespace jb {

// ids shared with include/janus_b200.h (JB200_HAM_*) and oracle/c/radau_oracle.c (PB_*)
enum HamId {
  HAM_HARMONIC = 0,
  HAM_VDP = 1,
  HAM_VDP_PMP_SMOOTH = 2,
  HAM_VDP_PMP_BANG = 3,
  HAM_VDP_AUG_BANG = 4,
  HAM_MLP = 5
};

// C1: H = (p^2 + w^2 x^2)/2    (written with squares like the reference's pow(2))
struct HamHarmonic {
  static constexpr int N = 1, NV = 2, NPAR = 1;
  static constexpr int FLOPS_F = 4, FLOPS_J = 2;
  static constexpr bool HAS_CONTROL = false;
  template <class P, class X, class PAR>
  JB_HD static auto H(const P& p, const X& x, const PAR& par, double) {
    const auto w = par[0];
    return (square(p) + (w * w) * square(x)) * 0.5;
  }
  JB_HD static double control(const double*, const double*) { return 0.0; }
};

// C2/C5: H = p1 x2 + p2 (mu (1 - x1^2) x2 - x1)
struct HamVdp {
  static constexpr int N = 2, NV = 4, NPAR = 1;
  static constexpr int FLOPS_F = 14, FLOPS_J = 16;
  static constexpr bool HAS_CONTROL = false;
  template <class P1, class P2, class X1, class X2, class PAR>
  JB_HD static auto H(const P1& p1, const P2& p2, const X1& x1, const X2& x2, const PAR& par,
                      double) {
    const auto mu = par[0];
    return p1 * x2 + p2 * (mu * (1.0 - x1 * x1) * x2 - x1);
  }
  JB_HD static double control(const double*, const double*) { return 0.0; }
};

// C3: Hu = p1 x2 + p2 u ((1 - x1^2) x2 - x1) + (W/2) u^2 + 1
//     (/root/reference/tests/cpp/ode_solver_test.cpp:185-197), u frozen.
template <bool BANG>
struct HamVdpPmp {
  static constexpr int N = 2, NV = 4, NPAR = 3;  // W, umin, umax
  static constexpr int FLOPS_F = BANG ? 16 : 34, FLOPS_J = BANG ? 18 : 38;
  static constexpr bool HAS_CONTROL = true;
  template <class P1, class P2, class X1, class X2, class PAR>
  JB_HD static auto H(const P1& p1, const P2& p2, const X1& x1, const X2& x2, const PAR& par,
                      double u) {
    const auto W = par[0];
    return p1 * x2 + p2 * u * ((1.0 - x1 * x1) * x2 - x1) + ((W * 0.5) * u * u + 1.0);   // W * 0.5 == W / 2 exactly
  }
  JB_HD static double control(const double* y, const double* par) {
    const double p2 = y[1], x1 = y[2], x2 = y[3];
    if (BANG)  // examples/cpp/radauted_vdp_aug_example.cpp:30-33
      return (p2 * (1.0 - x1 * x1) * x2 < 0.0) ? par[2] : par[1];
    return -p2 * ((1.0 - x1 * x1) * x2 - x1) / par[0];  // tests/cpp/ode_solver_test.cpp:17-24
  }
};

// "augmented" vdP of examples/cpp/radauted_vdp_aug_example.cpp:23-46 ==
// examples/python/aug_vdp_ref_test.py:6-27 :  H = p1 x2 + p2 (u (1 - x1^2) x2 - x1), bang-bang u
struct HamVdpAugBang {
  static constexpr int N = 2, NV = 4, NPAR = 2;  // umin, umax
  static constexpr int FLOPS_F = 16, FLOPS_J = 18;
  static constexpr bool HAS_CONTROL = true;
  template <class P1, class P2, class X1, class X2, class PAR>
  JB_HD static auto H(const P1& p1, const P2& p2, const X1& x1, const X2& x2, const PAR&,
                      double u) {
    return p1 * x2 + p2 * (u * (1.0 - x1 * x1) * x2 - x1);
  }
  JB_HD static double control(const double* y, const double* par) {
    const double p2 = y[1], x1 = y[2], x2 = y[3];
    return (p2 * (1.0 - x1 * x1) * x2 < 0.0) ? par[1] : par[0];
  }
};

// ------------------------------------------------------------------ adaptor
// parameters as the Hamiltonian sees them: plain doubles ...
struct ParPlain {
  const double* p;
  JB_HD double operator[](int k) const { return p[k]; }
};
// ... or, for the sensitivity with respect to parameter K (run-time index), hyper-dual numbers over NV + 1 variables whose
// extra variable NV is that parameter: par[K] is seeded with 1, every other parameter with an explicit 0
template <int NVX>
struct ParDual {
  const double* p;
  int K;
  JB_HD D2<NVX, (1u << (NVX - 1)), 0ull> operator[](int k) const {
    D2<NVX, (1u << (NVX - 1)), 0ull> r;
    r.v = p[k];
    r.g[NVX - 1] = (k == K) ? 1.0 : 0.0;
    return r;
  }
};

template <class Ham, size_t... I>
JB_HD auto ham_eval1(const double* y, const double* par, double u, std::index_sequence<I...>) {
  return Ham::H(seed1<Ham::NV, (int)I>(y[I])..., ParPlain{par}, u);
}
template <class Ham, size_t... I>
JB_HD auto ham_eval2(const double* y, const double* par, double u, std::index_sequence<I...>) {
  return Ham::H(seed2<Ham::NV, (int)I>(y[I])..., ParPlain{par}, u);
}

template <class Ham, size_t... I>
JB_HD auto ham_eval3(const double* y, const double* par, double u, std::index_sequence<I...>) {
  return Ham::H(seed3<Ham::NV, (int)I>(y[I])..., ParPlain{par}, u);
}
template <class Ham, size_t... I>
JB_HD auto ham_eval2p(const double* y, const double* par, int K, double u, std::index_sequence<I...>) {
  return Ham::H(seed2<Ham::NV + 1, (int)I>(y[I])..., ParDual<Ham::NV + 1>{par, K}, u);
}

// g = d f / d par[K] = [sigma d2H/dx dpar ; d2H/dp dpar]: the source term of a parameter direction of the forward
// sensitivities (the reference appends the parameter to the state instead, tests/cpp/ode_solver_test.cpp:630-640).
// A control law is evaluated on values and held constant (evalDynsU semantics), like in ham_jac.
template <class Ham>
JB_HD void ham_dpar(const double* y, const double* par, double sign, int K, double* g) {
  constexpr int N = Ham::N, NV = Ham::NV;
  const double u = Ham::HAS_CONTROL ? Ham::control(y, par) : 0.0;
  const auto r = ham_eval2p<Ham>(y, par, K, u, std::make_index_sequence<NV>{});
  using R = decltype(r);
#pragma unroll
  for (int rr = 0; rr < NV; rr++) {
    const int vi = (rr < N) ? (N + rr) : (rr - N);
    const int k = hidx(NV + 1, vi, NV);
    const double val = JB_BIT(R::hm, k) ? r.h[k] : 0.0;
    g[rr] = (rr < N) ? sign * val : val;
  }
}

// value + dual part of the dynamics and of the Jacobian (evalDynsDual / evalJacDual and their U twins,
// /root/reference/src/cpp/janus_ode_common.hpp:2075-2158, :2302-2416), from ONE third-order jet of H:
//   f [NV], J [NV*NV] as ham_dyns / ham_jac;  T [NV*NV*NV]: T[r][c][k] = d J[r][c] / d y_k  (the reference's eight
//   third-order blocks pxpppxH ... assembled in y-order, sigma on the pdot rows);
//   with ND directions yd [NV][ND]:  fd [NV][ND] = J yd,  Jd [NV][NV][ND] = T yd   (the .d parts of the reference's
//   TensorDual / TensorMatDual results).  Any of T, fd, Jd may be nullptr.
template <class Ham>
JB_HD void ham_jac_dual(const double* y, const double* par, double sign, const double* yd, int ND, double* f, double* J,
                        double* T, double* fd, double* Jd) {
  constexpr int N = Ham::N, NV = Ham::NV;
  const double u = Ham::HAS_CONTROL ? Ham::control(y, par) : 0.0;
  const D3<NV> r = ham_eval3<Ham>(y, par, u, std::make_index_sequence<NV>{});
  for (int rr = 0; rr < NV; rr++) {
    const int vi = (rr < N) ? (N + rr) : (rr - N);
    const double sg = (rr < N) ? sign : 1.0;
    f[rr] = sg * r.g[vi];
    for (int c = 0; c < NV; c++) {
      J[rr * NV + c] = sg * r.h[hsorted(NV, vi, c)];
      for (int k = 0; k < NV; k++) {
        const double t = sg * r.t[tsorted(NV, vi, c, k)];
        if (T) T[(rr * NV + c) * NV + k] = t;
      }
    }
  }
  if (!yd) return;
  for (int rr = 0; rr < NV; rr++) {
    const int vi = (rr < N) ? (N + rr) : (rr - N);
    const double sg = (rr < N) ? sign : 1.0;
    for (int l = 0; l < ND; l++) {
      if (fd) {
        double s = 0.0;
        for (int k = 0; k < NV; k++) s = fma(J[rr * NV + k], yd[k * ND + l], s);
        fd[rr * ND + l] = s;
      }
      if (Jd)
        for (int c = 0; c < NV; c++) {
          double s = 0.0;
          for (int k = 0; k < NV; k++) s = fma(sg * r.t[tsorted(NV, vi, c, k)], yd[k * ND + l], s);
          Jd[(rr * NV + c) * ND + l] = s;
        }
    }
  }
}

// f = [sigma dH/dx ; dH/dp]
template <class Ham>
JB_HD void ham_dyns(const double* y, const double* par, double sign, double* f) {
  constexpr int N = Ham::N, NV = Ham::NV;
  const double u = Ham::HAS_CONTROL ? Ham::control(y, par) : 0.0;
  const auto r = ham_eval1<Ham>(y, par, u, std::make_index_sequence<NV>{});
  using R = decltype(r);
#pragma unroll
  for (int i = 0; i < N; i++) {
    // pdot_i = sigma dH/dx_i (variable N+i) ; xdot_i = dH/dp_i (variable i)
    f[i] = JB_BIT(R::gm, N + i) ? sign * r.g[N + i] : 0.0;
    f[N + i] = JB_BIT(R::gm, i) ? r.g[i] : 0.0;
  }
}

// J[r][c] row-major NV x NV
template <class Ham>
JB_HD void ham_jac(const double* y, const double* par, double sign, double* J) {
  constexpr int N = Ham::N, NV = Ham::NV;
  const double u = Ham::HAS_CONTROL ? Ham::control(y, par) : 0.0;
  const auto r = ham_eval2<Ham>(y, par, u, std::make_index_sequence<NV>{});
  using R = decltype(r);
#pragma unroll
  for (int rr = 0; rr < NV; rr++) {
#pragma unroll
    for (int c = 0; c < NV; c++) {
      // ydot_rr = d H / d v, v = N+rr (rr<N) or rr-N ; then d/dy_c
      const int vi = (rr < N) ? (N + rr) : (rr - N);
      const int a = vi < c ? vi : c, b = vi < c ? c : vi;
      const int k = hidx(NV, a, b);
      double val = JB_BIT(R::hm, k) ? r.h[k] : 0.0;
      J[rr * NV + c] = (rr < N) ? sign * val : val;
    }
  }
}

}  // namespace jb
