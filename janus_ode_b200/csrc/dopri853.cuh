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
};

template <class Ham>
JB_HD void dopri_step_one(const double* par, double sign, long long norm_M, double atol, double rtol, bool aliased, double& x,
                          double* y, double* dydx, double& errold, double& hnext, double htry, double& hdid, int& nrej) {
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

}  // namespace dp
}  // namespace jb
