// Semi-implicit extrapolation stepper (SEULEX) for stiff batches, one thread = one
// trajectory; one kernel call = one step() of every trajectory of the batch.
//
// Replaces the reference's Python stepper /root/reference/src/python/seulext.py:
//   SeulexState :8-70   controller state (nseq 2,3,4,6,8,12,..., cost, coeff, k_targ, theta, jac_redo, flags)
//   SeulexT.step :158-581  order (k_targ) and step-size control around the extrapolation table
//   dy_batch :593-667   nseq[k] linearly implicit Euler steps  (I/h - df/dy) delta = f,  with the stability check
//                       theta = |delta_2| / max(1, |delta_1|) of the first two sub-steps for k <= 1
//   polyextr (src/python/seulex.py:439-462)  Aitken-Neville extrapolation in the sub-step count
// The reference is a masked-tensor translation of Numerical Recipes' StepperSie that loops over k and the
// sub-steps with one libtorch op per line for the whole batch (and builds / factors the matrix with its
// janus-linear LUBatch).  Here the whole step of one trajectory -- Jacobian by hyper-dual numbers, LU per table
// row, the Euler sequence, extrapolation, controller -- runs in one thread; the table lives in local memory.
// The Jacobian the reference keeps between steps (state.dfdy) is kept as the state it was evaluated at (yjac)
// and re-evaluated when needed: same bits, Ny instead of Ny^2 doubles per trajectory.
//
// Reference deviations from Numerical Recipes that are reproduced (S2-S5) or switchable (S1): see the header of
// the restatement oracle/c/radau_oracle.c (seulex_one), which this code equals bit for bit in mirror arithmetic.
#pragma once
#include "hamiltonians.cuh"
#include "lu.cuh"

namespace jb {
namespace sx {

constexpr int IMAXX_CAP = 16;   // KMAXX + 1 <= 16 (the reference examples use KMAXX = 12 or 13)
enum { F_FIRST = 1, F_LAST = 2, F_REJECT = 4, F_PREV_REJECT = 8 };
enum { Q_S1_CLOCK_ADVANCES_BY_HTRY = 1 };

struct StepArgs {
  long long M;
  double atol, rtol;
  int KMAXX, quirks;
  const double* htry;   // [M]
  double *y, *x;        // [M,NY], [M]      state, updated in place
  double *theta;        // [M]
  int* k_targ;          // [M]
  double *hnext, *hdid; // [M], [M]
  int* flags;           // [M]  F_*
  double* yjac;         // [M,NY] state the kept Jacobian was evaluated at
  int *k_out, *nrej;    // [M], [M]  (may be nullptr)
  // dual part (the reference's SeulexTed): yd [M,NY,nd] = d y / d (directions), updated in place across the step;
  // sjac [M,NY,nd] = the dual part of the state the kept Jacobian was evaluated at.  nullptr: none
  double* yd = nullptr;
  double* sjac = nullptr;
  int nd = 0;
};
constexpr int NDMAX = 8;   // dual directions per call

JB_HD int nseq_of(int i) {  // 2, 3, 4, 6, 8, 12, 16, 24, ...
  return (i & 1) ? (3 << (i >> 1)) : (2 << (i >> 1));
}
JB_HD double cost_of(int k) {  // costjac 5 + costlu 1 + nseq[0] (costfunc 1 + costsolv 1), then + (nseq[k+1]-1) 2 + 1
  double c = 5.0 + 1.0 + nseq_of(0) * (1.0 + 1.0);
  for (int j = 0; j < k; j++) c = c + (nseq_of(j + 1) - 1) * (1.0 + 1.0) + 1.0;
  return c;
}
JB_HD double coeff_of(int k, int l) { return 1.0 / ((double)nseq_of(k) / (double)nseq_of(l) - 1.0); }

template <class Ham>
JB_HD void seulex_step_one(const double* par, double sign, double atol, double rtol, int KMAXX, int quirks, double htry,
                           double* y, double& x_io, double& theta_io, int& ktarg_io, double& hnext_io, double& hdid_out,
                           int& flags_io, double* yjac, int& k_out, int& nrej_out, bool* jac_fresh = nullptr,
                           double* h_taken = nullptr) {
  constexpr int NY = Ham::NV;
  const double EPS_ = 2.220446049250313e-16;
  const int IMAXX = KMAXX + 1;
  double hopt[IMAXX_CAP + 1], work[IMAXX_CAP + 1], table[IMAXX_CAP][NY];
  double ysav[NY], yseq[NY], scale[NY], dfdy[NY * NY];
  const double jac_redo = fmin(1.0e-4, rtol);
  double x = x_io, theta = theta_io;
  int k_targ = ktarg_io;
  bool first_step = flags_io & F_FIRST, last_step = flags_io & F_LAST, reject = flags_io & F_REJECT,
       prev_reject = flags_io & F_PREV_REJECT, calcjac = false, have_jac = false;
  double errold = 0.0, fac = 0.0;
  for (int i = 0; i < IMAXX; i++) {
    hopt[i] = 0.0;
    work[i] = 0.0;
  }
  work[0] = 1.0e30;
  const double h = htry;
#pragma unroll
  for (int i = 0; i < NY; i++) ysav[i] = y[i];
  const bool forward = h > 0;
  if (h != hnext_io && !first_step) last_step = true;
  if (reject) {
    last_step = false;
    prev_reject = true;
    theta = 2.0 * jac_redo;
  }
#pragma unroll
  for (int i = 0; i < NY; i++) scale[i] = fma(rtol, fabs(y[i]), atol);
  reject = false;
  double hnew = fabs(h), h_att = fabs(h);
  int k = 0, nrej = 0;
  bool firstk = true;
  while (firstk || reject) {
    if (theta > jac_redo && !calcjac && fabs(h) > EPS_) {
#pragma unroll
      for (int i = 0; i < NY; i++) yjac[i] = y[i];
      ham_jac<Ham>(y, par, sign, dfdy);
      calcjac = true;
      have_jac = true;
    }
    if (!have_jac) {  // the Jacobian kept from an earlier step
      ham_jac<Ham>(yjac, par, sign, dfdy);
      have_jac = true;
    }
    firstk = false;
    reject = false;
    h_att = hnew;
    bool brk = false;
    k = 0;
    while (k <= k_targ + 1 && !brk) {
      bool success = true;
      {  // dy: nseq[k] linearly implicit Euler steps over h_att
        const int nsteps = nseq_of(k);
        const double hh = hnew / nsteps, rhh = 1.0 / hh;
        RLU<NY> lu;
#pragma unroll
        for (int i = 0; i < NY; i++)
#pragma unroll
          for (int j = 0; j < NY; j++) lu.a[i * NY + j] = (i == j) ? (-dfdy[i * NY + j] + rhh) : -dfdy[i * NY + j];
        rlu_factor<NY>(lu);
        double xnew = x + hh, ytemp[NY], delta[NY], yend[NY];
#pragma unroll
        for (int i = 0; i < NY; i++) ytemp[i] = ysav[i];
        ham_dyns<Ham>(ytemp, par, sign, delta);
        rlu_solve<NY>(lu, delta);
        for (int nn = 1; nn < nsteps; nn++) {
#pragma unroll
          for (int i = 0; i < NY; i++) ytemp[i] = ytemp[i] + delta[i];
          xnew += hh;
          ham_dyns<Ham>(ytemp, par, sign, yend);
          if (nn == 1 && k <= 1) {
            double del1 = 0.0, del2 = 0.0, dyt[NY];
#pragma unroll
            for (int i = 0; i < NY; i++) {
              const double v = delta[i] / scale[i];
              del1 = fma(v, v, del1);
            }
            del1 = ::sqrt(del1);
            ham_dyns<Ham>(ytemp, par, sign, dyt);
#pragma unroll
            for (int i = 0; i < NY; i++) delta[i] = dyt[i] - delta[i] * rhh;
            rlu_solve<NY>(lu, delta);
#pragma unroll
            for (int i = 0; i < NY; i++) {
              const double v = delta[i] / scale[i];
              del2 = fma(v, v, del2);
            }
            del2 = ::sqrt(del2);
            theta = del2 / fmax(1.0, del1);
            if (theta > 1.0) {
              success = false;
              break;
            }
          }
          rlu_solve<NY>(lu, yend);
#pragma unroll
          for (int i = 0; i < NY; i++) delta[i] = yend[i];
        }
        (void)xnew;   // autonomous Hamiltonian dynamics: the time argument is not used
#pragma unroll
        for (int i = 0; i < NY; i++) yseq[i] = ytemp[i] + delta[i];
      }
      if (!success) {
        hnew = fabs(hnew) * 0.5;
        reject = true;
        break;
      }
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < NY; i++) y[i] = yseq[i];
      } else {
#pragma unroll
        for (int i = 0; i < NY; i++) table[k - 1][i] = yseq[i];
        for (int j = k - 1; j > 0; j--) {  // polyextr
          const double c = coeff_of(k, j);
#pragma unroll
          for (int i = 0; i < NY; i++) table[j - 1][i] = fma(c, table[j][i] - table[j - 1][i], table[j][i]);
        }
        const double c0 = coeff_of(k, 0);
        double err = 0.0;
#pragma unroll
        for (int i = 0; i < NY; i++) {
          y[i] = fma(c0, table[0][i] - y[i], table[0][i]);
          scale[i] = fma(rtol, fabs(ysav[i]), atol);
          const double v = (y[i] - table[0][i]) / scale[i];
          err = fma(v, v, err);
        }
        err = ::sqrt(err / NY);
        if (err > 1.0 / EPS_ || (k > 1 && err >= errold)) {
          reject = true;
          hnew = fabs(hnew) * 0.5;
          break;
        }
        errold = fmax(4.0 * err, 1.0);
        const double expo = 1.0 / (k + 1);
        const double facmin = det_pow(0.1, expo);
        if (err == 0.0) {
          fac = 1.0 / facmin;
        } else {
          fac = 0.93 / det_pow(err / 0.6, expo);
          fac = fmax(facmin / 4.0, fmin(1.0 / facmin, fac));
        }
        hopt[k] = fabs(hnew * fac);
        work[k] = cost_of(k) / hopt[k];
        if ((first_step || last_step) && err <= 1.0) break;
        if (k == k_targ - 1 && !prev_reject && !first_step && !last_step) {
          if (err <= 1.0) break;
          if (err > (double)nseq_of(k_targ) * (double)nseq_of(k_targ + 1) * 4.0) {
            reject = true;
            k_targ = k;
            if (k_targ > 1 && work[k - 1] < 0.7 * work[k]) {  // S4
              k_targ -= 1;
              hnew = hopt[k_targ];
              break;
            }
          }
        }
        if (k == k_targ) {
          if (err <= 1.0) break;
          if (err > nseq_of(k + 1) * 2.0) {
            reject = true;
            if (k_targ > 1 && work[k - 1] < 0.7 * work[k]) k_targ -= 1;
            hnew = hopt[k_targ];
            break;
          }
        }
        if (k == k_targ + 1) {
          if (err > 1.0) {
            reject = true;
            if (k_targ > 1 && work[k_targ - 1] < 0.7 * work[k_targ]) k_targ -= 1;
            hnew = hopt[k_targ];
          }
          break;
        }
      }
      k += 1;
    }
    (void)brk;
    if (reject) {
      nrej += 1;
      prev_reject = true;
      if (!calcjac) theta = theta * jac_redo;  // S2
      if (nrej > 200) break;
    }
  }
  const double hs = (quirks & Q_S1_CLOCK_ADVANCES_BY_HTRY) ? h : (forward ? h_att : -h_att);
  x = x + hs;
  hdid_out = hs;
  first_step = false;
  int kopt = k;
  if (k == 1) kopt = 2;
  if (k <= k_targ) {
    kopt = k;
    if (work[k - 1] < work[k] * 0.7) kopt = k - 1;
    if (work[k] < 0.9 * work[k - 1]) kopt = (k + 1 < KMAXX) ? k + 1 : KMAXX;  // S3
  }
  if (k != 1 && k > k_targ) {
    kopt = k - 1;
    if (k > 2 && work[k - 2] < 0.7 * work[k - 1]) kopt = k - 2;
    if (work[k] < 0.9 * work[kopt]) kopt = (k < KMAXX) ? k : KMAXX;
  }
  if (prev_reject) {
    k_targ = kopt < k ? kopt : k;
    hnew = fmin(fabs(hs), hopt[k_targ]);
    prev_reject = false;
  } else {
    if (kopt <= k) hnew = hopt[kopt];
    else if (k < k_targ && work[k] < 0.9 * work[k - 1]) hnew = hopt[k] * cost_of(kopt + 1 < KMAXX ? kopt + 1 : KMAXX) / cost_of(k);
    else hnew = hopt[k] * cost_of(kopt) / cost_of(k);
    k_targ = kopt;
  }
  hnext_io = forward ? hnew : -hnew;
  x_io = x;
  theta_io = theta;
  ktarg_io = k_targ;
  flags_io = (first_step ? F_FIRST : 0) | (last_step ? F_LAST : 0) | (reject ? F_REJECT : 0) | (prev_reject ? F_PREV_REJECT : 0);
  k_out = k;
  nrej_out = nrej;
  if (jac_fresh) *jac_fresh = calcjac;
  if (h_taken) *h_taken = forward ? h_att : -h_att;
}

// Dual part of one ACCEPTED step (the reference's SeulexTed, /root/reference/src/python/seulexted.py:238-1396: TensorDual
// state, TensorMatDual Jacobian and LUDual solves through dy / polyextr).  The accepted attempt -- table rows 0..k over
// h_att from ysav with the iteration matrix A = J(yjac) -- is replayed on [y, y']:
//   (I / hh - A) delta = f(y_n)   =>   (I / hh - A) delta' = J(y_n) y_n' + A' delta,     A' = dJ/dy (yjac) . Sjac,
// A' from the third-order jet of H (dual3.cuh, the reference's evalJacDual), J(y_n) from hyper-dual numbers, the same
// real LU; the Aitken-Neville extrapolation is linear, so the dual table is extrapolated with the same coefficients.
// Order and step-size control look at the real part only: the states are those of seulex_step_one bit for bit.
// Sd [NY][nd] (in: dual part at the start of the step, out: at its end), Sj [NY][nd] dual part of yjac; nd <= NDMAX.
template <class Ham>
JB_HD void seulex_dual_one(const double* par, double sign, const double* ysav, const double* yjac, const double* Sj, double h_att,
                           int k, double* Sd, int nd) {
  constexpr int NY = Ham::NV;
  double A[NY * NY], Ad[NY * NY * NDMAX], f0[NY];
  {
    double Jd[NY * NY * NDMAX], yd[NY * NDMAX];
    for (int i = 0; i < NY; i++)
      for (int d = 0; d < nd; d++) yd[i * nd + d] = Sj[i * nd + d];
    ham_jac_dual<Ham>(yjac, par, sign, yd, nd, f0, A, nullptr, nullptr, Jd);
    for (int e = 0; e < NY * NY; e++)
      for (int d = 0; d < nd; d++) Ad[e * NDMAX + d] = Jd[e * nd + d];
  }
  double tabd[IMAXX_CAP][NY * NDMAX], outd[NY * NDMAX];
  for (int kk = 0; kk <= k; kk++) {
    const int nsteps = nseq_of(kk);
    const double hh = h_att / nsteps, rhh = 1.0 / hh;
    RLU<NY> lu;
    for (int i = 0; i < NY; i++)
      for (int j = 0; j < NY; j++) lu.a[i * NY + j] = (i == j) ? (-A[i * NY + j] + rhh) : -A[i * NY + j];
    rlu_factor<NY>(lu);
    double yt[NY], ytd[NY * NDMAX], delta[NY], deltad[NY * NDMAX], J[NY * NY];
    for (int i = 0; i < NY; i++) {
      yt[i] = ysav[i];
      for (int d = 0; d < nd; d++) ytd[i * NDMAX + d] = Sd[i * nd + d];
    }
    for (int nn = 0; nn < nsteps; nn++) {
      if (nn > 0)
        for (int i = 0; i < NY; i++) {
          yt[i] = yt[i] + delta[i];
          for (int d = 0; d < nd; d++) ytd[i * NDMAX + d] += deltad[i * NDMAX + d];
        }
      ham_dyns<Ham>(yt, par, sign, delta);
      rlu_solve<NY>(lu, delta);
      ham_jac<Ham>(yt, par, sign, J);
      for (int d = 0; d < nd; d++) {
        double rhs[NY];
        for (int i = 0; i < NY; i++) {
          double acc = 0.0;
          for (int j = 0; j < NY; j++) acc = fma(J[i * NY + j], ytd[j * NDMAX + d], acc);
          for (int j = 0; j < NY; j++) acc = fma(Ad[(i * NY + j) * NDMAX + d], delta[j], acc);
          rhs[i] = acc;
        }
        rlu_solve<NY>(lu, rhs);
        for (int i = 0; i < NY; i++) deltad[i * NDMAX + d] = rhs[i];
      }
    }
    double seqd[NY * NDMAX];
    for (int i = 0; i < NY; i++)
      for (int d = 0; d < nd; d++) seqd[i * NDMAX + d] = ytd[i * NDMAX + d] + deltad[i * NDMAX + d];
    if (kk == 0) {
      for (int e = 0; e < NY * NDMAX; e++) outd[e] = seqd[e];
    } else {
      for (int e = 0; e < NY * NDMAX; e++) tabd[kk - 1][e] = seqd[e];
      for (int j = kk - 1; j > 0; j--) {
        const double c = coeff_of(kk, j);
        for (int i = 0; i < NY; i++)
          for (int d = 0; d < nd; d++) {
            const int e = i * NDMAX + d;
            tabd[j - 1][e] = fma(c, tabd[j][e] - tabd[j - 1][e], tabd[j][e]);
          }
      }
      const double c0 = coeff_of(kk, 0);
      for (int i = 0; i < NY; i++)
        for (int d = 0; d < nd; d++) {
          const int e = i * NDMAX + d;
          outd[e] = fma(c0, tabd[0][e] - outd[e], tabd[0][e]);
        }
    }
  }
  for (int i = 0; i < NY; i++)
    for (int d = 0; d < nd; d++) Sd[i * nd + d] = outd[i * NDMAX + d];
}

}  // namespace sx
}  // namespace jb
