// Explicit Dormand-Prince 8(5,3) stepper for non-stiff batches, one thread = one
// trajectory, the whole attempted-step loop of one call in registers.
//
// Mirrors the reference's Python stepper /root/reference/src/python/dopri853te.py (the only reference
// solver that can be executed in the build container, so this path is pinned against live reference
// output: tests/golden/make_dopri_golden.py):
//   dy()      :207-287  12 stages, 8th-order solution, the two error vectors yerr / yerr2
//   error()   :191-205  err = |h| * S2 * sqrt(1 / (D * M * (S2 + 0.01 * S1))),  S1 = sum (yerr/sk)^2,
//                       S2 = sum (yerr2/sk)^2, sk = atol + rtol * max(|y|, |yout|)
//   control() :318-355  accept if err <= 1
//   step()    :290-315  retry with h *= scale until accepted, then dydx = f(x + h, yout)
// Reference quirks reproduced (they change the step sequence, so parity needs them):
//   * every tableau coefficient and `safe`, `minscale` are float32-rounded (torch.tensor(<python float>));
//   * the error norm divides by D * M with M = the BATCH size (:205), so tolerances tighten with the batch;
//     `norm_M` carries that number (a sharded run passes the global batch size);
//   * control() starts from reject = all-true (:320), so hnext = h * min(scale, 1): the proposed step never grows;
//   * beta = 0: errold^beta = 1 (errold is still tracked: scale = maxscale when it is exactly 0);
//   * step() ends with `self.y = self.yout` (:311): from the second call on y and yout are ONE tensor, so the
//     attempt's result overwrites y before error() reads it (sk = atol + rtol |yout|) and a rejected attempt
//     leaves its yout behind as the state the retry starts from, with the old dydx (`aliased`).
// err^(-1/8) is evaluated as 1 / sqrt(sqrt(sqrt(err))) (three correctly-rounded square roots: the same bits
// on the GPU and in the numpy oracle; the reference calls torch.pow).
#pragma once
#include "dopri853_tables.cuh"
#include "hamiltonians.cuh"

namespace jb {
namespace dp {

struct StepArgs {
  long long M, norm_M;
  double atol, rtol;
  double *x, *y, *dydx, *errold, *hnext;   // state, updated in place ([M], [M,NY], [M,NY], [M], [M])
  int aliased;                              // 0: first step() of this stepper object, 1: later calls
  const double* htry;                       // [M]
  double* hdid;                             // [M]
  int* nrej;                                // [M] rejected attempts of this call (may be nullptr)
  double* yd = nullptr;                     // [M,NY,nd] dual part of y (forward sensitivities), updated in place; nullptr: none
  int nd = 0;
};
constexpr int NDMAX = 8;                    // dual directions per call

template <class Ham>
JB_HD void dopri_step_one(const double* par, double sign, long long norm_M, double atol, double rtol, bool aliased, double& x,
                          double* y, double* dydx, double& errold, double& hnext, double htry, double& hdid, int& nrej,
                          double* ybase = nullptr) {
  constexpr int NY = Ham::NV;
  const double EPS_ = 2.220446049250313e-16;
  double h = htry;
  nrej = 0;
  double yout[NY];
#pragma unroll
  for (int i = 0; i < NY; i++) yout[i] = y[i];
  bool reject = h > EPS_;
  while (reject) {
    double k4[NY], k5[NY], k6[NY], k7[NY], k8[NY], k9[NY], k10[NY], k2[NY], k3[NY], yt[NY];
    if (ybase) {   // the state this attempt starts from (the dual pass repeats the accepted attempt)
#pragma unroll
      for (int i = 0; i < NY; i++) ybase[i] = y[i];
    }
#pragma unroll
    for (int i = 0; i < NY; i++) yt[i] = y[i] + h * (A2_1 * dydx[i]);
    ham_dyns<Ham>(yt, par, sign, k2);
#pragma unroll
    for (int i = 0; i < NY; i++) yt[i] = y[i] + h * (A3_1 * dydx[i] + A3_2 * k2[i]);
    ham_dyns<Ham>(yt, par, sign, k3);
#pragma unroll
    for (int i = 0; i < NY; i++) yt[i] = y[i] + h * (A4_1 * dydx[i] + A4_3 * k3[i]);
    ham_dyns<Ham>(yt, par, sign, k4);
#pragma unroll
    for (int i = 0; i < NY; i++) yt[i] = y[i] + h * (A5_1 * dydx[i] + A5_3 * k3[i] + A5_4 * k4[i]);
    ham_dyns<Ham>(yt, par, sign, k5);
#pragma unroll
    for (int i = 0; i < NY; i++) yt[i] = y[i] + h * (A6_1 * dydx[i] + A6_4 * k4[i] + A6_5 * k5[i]);
    ham_dyns<Ham>(yt, par, sign, k6);
#pragma unroll
    for (int i = 0; i < NY; i++) yt[i] = y[i] + h * (A7_1 * dydx[i] + A7_4 * k4[i] + A7_5 * k5[i] + A7_6 * k6[i]);
    ham_dyns<Ham>(yt, par, sign, k7);
#pragma unroll
    for (int i = 0; i < NY; i++)
      yt[i] = y[i] + h * (A8_1 * dydx[i] + A8_4 * k4[i] + A8_5 * k5[i] + A8_6 * k6[i] + A8_7 * k7[i]);
    ham_dyns<Ham>(yt, par, sign, k8);
#pragma unroll
    for (int i = 0; i < NY; i++)
      yt[i] = y[i] + h * (A9_1 * dydx[i] + A9_4 * k4[i] + A9_5 * k5[i] + A9_6 * k6[i] + A9_7 * k7[i] + A9_8 * k8[i]);
    ham_dyns<Ham>(yt, par, sign, k9);
#pragma unroll
    for (int i = 0; i < NY; i++)
      yt[i] = y[i] + h * (A10_1 * dydx[i] + A10_4 * k4[i] + A10_5 * k5[i] + A10_6 * k6[i] + A10_7 * k7[i] + A10_8 * k8[i] +
                          A10_9 * k9[i]);
    ham_dyns<Ham>(yt, par, sign, k10);
#pragma unroll
    for (int i = 0; i < NY; i++)
      yt[i] = y[i] + h * (A11_1 * dydx[i] + A11_4 * k4[i] + A11_5 * k5[i] + A11_6 * k6[i] + A11_7 * k7[i] + A11_8 * k8[i] +
                          A11_9 * k9[i] + A11_10 * k10[i]);
    ham_dyns<Ham>(yt, par, sign, k2);   // stage 11 (the reference reuses k2, :266)
#pragma unroll
    for (int i = 0; i < NY; i++)
      yt[i] = y[i] + h * (A12_1 * dydx[i] + A12_4 * k4[i] + A12_5 * k5[i] + A12_6 * k6[i] + A12_7 * k7[i] + A12_8 * k8[i] +
                          A12_9 * k9[i] + A12_10 * k10[i] + A12_11 * k2[i]);
    ham_dyns<Ham>(yt, par, sign, k3);   // stage 12 (reuses k3, :273)
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int i = 0; i < NY; i++) {
      const double kk = B1 * dydx[i] + B6 * k6[i] + B7 * k7[i] + B8 * k8[i] + B9 * k9[i] + B10 * k10[i] + B11 * k2[i] +
                        B12 * k3[i];
      yout[i] = y[i] + h * kk;
      const double e1 = kk - BHH1 * dydx[i] - BHH2 * k9[i] - BHH3 * k3[i];
      const double e2 = k5[i] - BHH1 * dydx[i] - BHH2 * k9[i] - BHH3 * k3[i];
      const double sk = atol + rtol * fmax(fabs(aliased ? yout[i] : y[i]), fabs(yout[i]));
      const double q1 = e1 / sk, q2 = e2 / sk;
      s1 = s1 + q1 * q1;
      s2 = s2 + q2 * q2;
    }
    double deno = s2 + 0.01 * s1;
    if (deno <= 0.0) deno = 1.0;
    const double err = fabs(h) * s2 * ::sqrt(1.0 / ((double)(NY * norm_M) * deno));
    // control (:318-355)
    if (err <= 1.0) {
      double scale;
      if (errold == 0.0) {
        scale = MAXSCALE;
      } else {
        scale = SAFE * (1.0 / ::sqrt(::sqrt(::sqrt(err))));   // err^-alpha * errold^beta, alpha = 1/8, beta = 0
        if (scale < MINSCALE) scale = MINSCALE;
        if (scale > MAXSCALE) scale = MAXSCALE;
      }
      hnext = h * fmin(scale, 1.0);
      errold = fmax(err, 1.0e-4);
      reject = false;
    } else if (err > 1.0) {
      const double scale = fmax(SAFE * (1.0 / ::sqrt(::sqrt(::sqrt(err)))), MINSCALE);
      h = h * scale;
      nrej += 1;
      if (aliased) {
#pragma unroll
        for (int i = 0; i < NY; i++) y[i] = yout[i];
      }
    } else {   // NaN: the reference would loop forever; give up on this trajectory
      h = 0.0 / 0.0;
      reject = false;
    }
  }
  if (h == h && htry > EPS_) {
    double yn[NY];
#pragma unroll
    for (int i = 0; i < NY; i++) yn[i] = yout[i];
    ham_dyns<Ham>(yn, par, sign, dydx);
#pragma unroll
    for (int i = 0; i < NY; i++) y[i] = yout[i];
    x = x + h;
  }
  hdid = h;
}

// Dual part of one ACCEPTED step (the reference's Dopri853TeD, /root/reference/src/cpp/dopri853ted.hpp:26-734: TensorDual
// state through dy()): an explicit Runge-Kutta step differentiates stage by stage,
//   Y_s' = S + h sum_j a_sj K_j',   K_s' = J(Y_s) Y_s',   S_new = S + h sum_j b_j K_j',
// J(Y_s) from the hyper-dual evaluation of H at the stage points (which are recomputed here from the step's base state, its
// derivative k1 and the accepted h).  Step size and error control see the real part only, so the states are those of
// dopri_step_one bit for bit.  Sd [NY][nd] row-major, nd <= NDMAX.
template <class Ham>
JB_HD void dopri_dual_one(const double* par, double sign, const double* yb, const double* k1, double h, double* Sd, int nd) {
  constexpr int NY = Ham::NV;
  const double A[11][11] = {
      {A2_1},
      {A3_1, A3_2},
      {A4_1, 0, A4_3},
      {A5_1, 0, A5_3, A5_4},
      {A6_1, 0, 0, A6_4, A6_5},
      {A7_1, 0, 0, A7_4, A7_5, A7_6},
      {A8_1, 0, 0, A8_4, A8_5, A8_6, A8_7},
      {A9_1, 0, 0, A9_4, A9_5, A9_6, A9_7, A9_8},
      {A10_1, 0, 0, A10_4, A10_5, A10_6, A10_7, A10_8, A10_9},
      {A11_1, 0, 0, A11_4, A11_5, A11_6, A11_7, A11_8, A11_9, A11_10},
      {A12_1, 0, 0, A12_4, A12_5, A12_6, A12_7, A12_8, A12_9, A12_10, A12_11}};
  const double B[12] = {B1, 0, 0, 0, 0, B6, B7, B8, B9, B10, B11, B12};
  double K[12][NY], Kd[12][NY * NDMAX], J[NY * NY];
  for (int i = 0; i < NY; i++) K[0][i] = k1[i];
  ham_jac<Ham>(yb, par, sign, J);
  for (int i = 0; i < NY; i++)
    for (int d = 0; d < nd; d++) {
      double acc = 0.0;
      for (int j = 0; j < NY; j++) acc = fma(J[i * NY + j], Sd[j * nd + d], acc);
      Kd[0][i * NDMAX + d] = acc;
    }
  for (int s = 1; s < 12; s++) {
    double Y[NY], Yd[NY * NDMAX];
    for (int i = 0; i < NY; i++) {
      double acc = 0.0;
      for (int j = 0; j < s; j++) acc = fma(A[s - 1][j], K[j][i], acc);
      Y[i] = fma(h, acc, yb[i]);
      for (int d = 0; d < nd; d++) {
        double ad = 0.0;
        for (int j = 0; j < s; j++) ad = fma(A[s - 1][j], Kd[j][i * NDMAX + d], ad);
        Yd[i * NDMAX + d] = fma(h, ad, Sd[i * nd + d]);
      }
    }
    ham_dyns<Ham>(Y, par, sign, K[s]);
    ham_jac<Ham>(Y, par, sign, J);
    for (int i = 0; i < NY; i++)
      for (int d = 0; d < nd; d++) {
        double acc = 0.0;
        for (int j = 0; j < NY; j++) acc = fma(J[i * NY + j], Yd[j * NDMAX + d], acc);
        Kd[s][i * NDMAX + d] = acc;
      }
  }
  for (int i = 0; i < NY; i++)
    for (int d = 0; d < nd; d++) {
      double acc = 0.0;
      for (int j = 0; j < 12; j++) acc = fma(B[j], Kd[j][i * NDMAX + d], acc);
      Sd[i * nd + d] = fma(h, acc, Sd[i * nd + d]);
    }
}

}  // namespace dp
}  // namespace jb
