// Per-trajectory variable-order Radau IIA stepper (1/3/5/7 stages), one thread = one trajectory.
//
// B200-native replacement of the reference's masked batch loop
//   /root/reference/src/cpp/radaute_impl.hpp:537-1345 (RadauTe::solve), :1502-1661 (DecomRC),
//   :1721-1897 (Solvrad), :1902-1986 (Estrad); scalar truth src/matlab/radau.m:630-1256;
//   per-trajectory (any()) semantics of src/cpp/radauted_impl.hpp:620-1645.
// The reference runs ~200 masked libtorch ops + ~100 host syncs per pass over [M,...] tensors.
// Here the whole state machine of one trajectory -- Hamiltonian derivatives by (hyper-)dual
// numbers, simplified Newton, real + complex LU, error estimate, step-size / order controller
// -- runs inside one thread with its state in registers / local memory; HBM is touched once
// per trajectory (load y0, tspan; store y, t, h, res, stats).
//
// Everything here is __host__ __device__ so that tests/hostsim can run the very same code on
// the CPU for debugging (never part of the product library).
//
// Reference deviations kept/dropped: SURVEY.md 8a Q3, Q4, Q6-Q9 (same as oracle/radau_scalar.py).
// ARITHMETIC CONTRACT: every floating-point operation below is a fixed IEEE sequence (explicit
// fma() where fused, plain ops otherwise; the translation unit is compiled with -fmad=false so
// nothing else is contracted).  oracle/c/radau_oracle.c in "mirror" mode follows the same
// sequence, which gives bit-identical trajectories and counters.
#pragma once
#include "hamiltonians.cuh"
#include "lu.cuh"
#include "radau_tables_t.cuh"

namespace jb {

constexpr int NYMAX = 16;
constexpr double EPS = 2.220446049250313e-16;

enum { ST_FCN, ST_JAC, ST_DEC, ST_SOL, ST_STEP, ST_ACC, ST_REJ, ST_NREJ };

// per-solve constants, filled on the host by fill_params() (csrc/capi.cu)
struct RadauParams {
  double RelTol1[4][NYMAX];  // 0.1 RelTol^ExpmNs        per stage count (index (s-1)/2)
  double AbsTol1[4][NYMAX];  // RelTol1 * AbsTol/RelTol
  double FNewt[4];           // Newton stopping threshold (radau.m:897-900)
  double inv_sqrt_sn[4];     // 1/sqrt(s*Ny)
  double inv_sqrt_n;         // 1/sqrt(Ny)
  double InitialStep, Thet, Safe, Quot1, Quot2, FacL, FacR;  // FacL, FacR already inverted
  double Vitu, Vitd, hhou, hhod;
  double sign;
  long long MaxNbrStep;
  int Start_Newt, MaxNbrNewton, NbrStg, MinNbrStg, MaxNbrStg;
  int Gustafsson;            // bit 0: on; bit 1: limiter not gated by ChangeFlag (RadauTe, quirk Q8)
  double ham_par[8];
  // event functions E_k(t, y) = ev_A[k] . y - ev_b[k]  (EventsFcn of the reference: value, isterminal, direction --
  // src/cpp/radaute.hpp:44-47; zero location src/matlab/radau.m:1494-1618).  ev_K = 0: none.
  int key_fine;              // scheduling only: epoch lists also ordered by the current step size (solver_impl.cuh regime_key)
  int ev_K, ev_stop[4], ev_slope[4];
  double ev_A[4][NYMAX], ev_b[4];
  // forward sensitivities: parameter directions d y / d ham_par[sens_par[k]], k < sens_npar (sens_step)
  int sens_npar, sens_par[4];
};
#define JB_SENS_MAXPAR 4

// Algorithmic FLOP model of SURVEY.md 8(d) (mul, add, div, sqrt, pow = 1; complex mul = 6, add = 2),
// n = Ny, s stages, F / Jf = FP64 instructions of one dynamics / Jacobian evaluation counted from
// this library's own SASS (Ham::FLOPS_F, Ham::FLOPS_J).  Accumulated per trajectory in Traj::work.
JB_HD constexpr long long W_newt(int s, int n, int F) {  // one simplified-Newton iteration
  return (long long)s * F + 2LL * s * s * n + 2LL * s * s * n + (long long)s * n + (2LL * n + 8LL * n * ((s - 1) / 2)) +
         (2LL * n * n + 8LL * n * n * ((s - 1) / 2)) + 3LL * s * n;
}
JB_HD constexpr long long W_lu(int s, int n) {  // real + (s-1)/2 complex LU factorisations
  return (2LL * n * n * n + 8LL * n * n * n * ((s - 1) / 2)) / 3;
}
JB_HD constexpr long long W_tail(int s, int n, int F) {  // Estrad + f0 + controller
  return (long long)F + 2LL * s * n + 2LL * n * n + 3LL * n + 60;
}

JB_HD constexpr int nit_of(int s, int MaxNbrNewton) {
  return s == 1 ? MaxNbrNewton - 3 : s == 3 ? MaxNbrNewton : s == 5 ? MaxNbrNewton + 5 : MaxNbrNewton + 10;
}

// Output at the interior points of tspan (OutFlag 3, /root/reference/src/matlab/radau.m:1150-1168;
// the C++ reference leaves this case as a TODO, src/cpp/radaute_impl.hpp:1206-1207): rows of ONE
// trajectory.  tspan == nullptr: final state only.
struct DenseOut {
  const double* tspan;  // [ntspan] requested output times (first = t0, last = tfinal)
  int ntspan;
  double* tout;         // [ntspan]   (mode 1 / 2: [cap])
  double* yout;         // [ntspan][NY]
  // mode 3 (default): output at the tspan points.  mode 1: every accepted step (OutFlag 1, radaute_impl.hpp:1166-1183);
  // mode 2: plus refine - 1 interpolated points per step (OutFlag 2, radau.m:1130-1143).  cap = rows of tout / yout:
  // the reference grows its buffers (torch::cat of the whole [M, .] tensor on every overflow); here the caller sizes
  // them, points beyond cap are counted in nout but not stored.
  int mode = 3, cap = 0, refine = 1;
  // located events of this trajectory (radau.m:1110-1121): te [ecap], ye [ecap][NY], ie [ecap]; *nev counts them
  double* te = nullptr;
  double* ye = nullptr;
  int* ie = nullptr;
  int* nev = nullptr;
  int ecap = 0;
};

template <int NY, int SM>
struct Traj {
  // The Jacobian itself is NOT part of the persistent state: it is a deterministic function of the state it was
  // evaluated at (yJ), so a factorisation that needs a Jacobian older than this pass re-evaluates it (same code,
  // same bits).  That keeps Ny^2 doubles out of the registers across the Newton loop and out of the parked record.
  double y[NY], f0[NY], yJ[NY], cont[SM * NY];
  double t, tfinal, h, h_old, hopt, hacc, erracc, FacConv, Theta, Thetat, PosNeg, hmaxn;
  int NbrStg, Newt, ChangeNbr, nout;
  bool Last, First, Reject, Keep, ChangeFlag, UnExpStepRej, UnExpNewtRej, NeedNewJac, NeedNewLU;
  int res;
  int st[8];
  long long work;  // algorithmic FLOPs spent on this trajectory (SURVEY.md 8d formula, see W_* below)
  RLU<NY> L0;
  // complex blocks q = 1..(SM-1)/2: the first LcPlace::NS in shared memory (lu.cuh ZLUS), the rest here
  using Place = LcPlace<NY, SM>;
  ZLUS<NY> LcS[Place::NS > 0 ? Place::NS : 1];
  ZLU<NY> Lc[Place::NR > 0 ? Place::NR : 1];
#if !defined(__CUDA_ARCH__)
  ZLU<NY> hiblk[3];   // host pass of the rolled high-order stepper (device: shared memory, lu.cuh ZLUC)
#endif
};

// complex block `block` (run-time index) of the rolled high-order stepper
template <int NY, int SM>
JB_HD void hi_store(Traj<NY, SM>& T, int block, const ZLU<NY>& Z) {
#if defined(__CUDA_ARCH__) && defined(JB_LC_SMEM)
  zluc_store<NY>(block, Z);
#else
  T.hiblk[block] = Z;
#endif
}
template <int NY, int SM>
JB_HD void hi_solve(const Traj<NY, SM>& T, int block, double* br, double* bi) {
#if defined(__CUDA_ARCH__) && defined(JB_LC_SMEM)
  zluc_solve<NY>(block, br, bi);
#else
  zlu_solve<NY>(T.hiblk[block], br, bi);
#endif
}

JB_HD int imin(int a, int b) { return a < b ? a : b; }
JB_HD int imax(int a, int b) { return a > b ? a : b; }
JB_HD bool bad(double x) { return !(x - x == 0.0); }  // NaN or Inf

JB_HD double ipow(double x, int n) {
  double p = 1.0;
  for (int k = 0; k < n; k++) p = p * x;
  return p;
}

template <class Ham, int NY, int SM>
JB_HD void traj_init(Traj<NY, SM>& T, const RadauParams& P, const double* y0, double t0, double tfinal) {
  // radau.m:631-696 / radaute_impl.hpp:406-455
#pragma unroll
  for (int i = 0; i < NY; i++) T.y[i] = T.yJ[i] = y0[i];
#pragma unroll
  for (int i = 0; i < 8; i++) T.st[i] = 0;
#pragma unroll
  for (int i = 0; i < SM * NY; i++) T.cont[i] = 0.0;
  T.t = t0;
  T.tfinal = tfinal;
  T.PosNeg = (tfinal - t0) >= 0 ? 1.0 : -1.0;
  T.NbrStg = P.NbrStg;
  T.nout = 0;
  T.UnExpStepRej = T.UnExpNewtRej = T.Keep = T.ChangeFlag = false;
  T.ChangeNbr = 0;
  T.Theta = T.Thetat = 0.0;
  T.hmaxn = fabs(tfinal - t0);
  double h = P.InitialStep;
  if (fabs(h) <= 10 * EPS) h = 1e-6;
  h = T.PosNeg * fmin(fabs(h), T.hmaxn);
  T.h_old = h;
  T.hopt = h;
  if ((T.t + h * 1.0001 - tfinal) * T.PosNeg >= 0) {
    h = tfinal - T.t;
    T.Last = true;
  } else {
    T.Last = false;
  }
  T.h = h;
  T.FacConv = 1.0;
  T.Reject = false;
  T.First = true;
  ham_dyns<Ham>(T.y, P.ham_par, P.sign, T.f0);
  T.st[ST_FCN] += 1;
  T.work = Ham::FLOPS_F;
  T.res = 0;
#pragma unroll
  for (int i = 0; i < NY; i++)
    if (T.f0[i] != T.f0[i]) T.res = 2;
  T.NeedNewJac = true;
  T.NeedNewLU = true;
  T.Newt = 0;
  T.hacc = 0.0;
  T.erracc = 0.0;
}

// gamma/h I - J and (alpha_q + i beta_q)/h I - J, factored (radaute_impl.hpp:1502-1625)
template <int NY, int S, int SM>
JB_HD void decomrc(Traj<NY, SM>& T, const double* Jac) {
  using Tab = RadauTab<S>;
  const double rh = 1.0 / T.h;
  const double valp0 = Tab::ValP(0) * rh;
#pragma unroll
  for (int i = 0; i < NY; i++)
#pragma unroll
    for (int j = 0; j < NY; j++) T.L0.a[i * NY + j] = (i == j) ? (valp0 - Jac[i * NY + j]) : -Jac[i * NY + j];
  rlu_factor<NY>(T.L0);
#pragma unroll
  for (int q = 1; q <= (S - 1) / 2; q++) {
    const double al = Tab::ValP(2 * q - 1) * rh, be = Tab::ValP(2 * q) * rh;
    constexpr int NS = LcPlace<NY, SM>::NS;
    const bool parked = (q - 1) < NS;
    ZLU<NY> Ztmp;
    ZLU<NY>& Z = parked ? Ztmp : T.Lc[(q - 1 - NS) > 0 ? (q - 1 - NS) : 0];
#pragma unroll
    for (int i = 0; i < NY; i++)
#pragma unroll
      for (int j = 0; j < NY; j++) {
        Z.re[i * NY + j] = (i == j) ? (al - Jac[i * NY + j]) : -Jac[i * NY + j];
        Z.im[i * NY + j] = (i == j) ? be : 0.0;
      }
    zlu_factor<NY>(Z);
    if (parked) zlus_store<NY>(T.LcS[(q - 1) < NS ? (q - 1) : 0], Ztmp, q - 1);
  }
}

// one pass of the outer loop = one attempted step (radau.m:733-1253).
// returns true when the trajectory is finished (done or failed; T.res holds the code)
// dense output after an accepted step: every requested time in (t - h, t] is evaluated on the
// collocation polynomial (ntrprad, src/cpp/radaute_impl.hpp:1413-1440 == radau.m:1291-1308)
// collocation polynomial of the step that just ended at (T.t, T.y) with size T.h_old, evaluated at time tn
template <int NY, int S, int SM>
JB_HD void ntrprad(const Traj<NY, SM>& T, double tn, double* yi_out) {
  using Tab = RadauTab<S>;
  const double s = (tn - T.t) / T.h_old;
#pragma unroll
  for (int i = 0; i < NY; i++) {
    double yi = (s - (Tab::C(0) - 1.0)) * T.cont[(S - 1) * NY + i];
#pragma unroll
    for (int q = 1; q < S; q++) yi = (s - (Tab::C(q) - 1.0)) * (yi + T.cont[(S - 1 - q) * NY + i]);
    yi_out[i] = yi + T.y[i];
  }
}

JB_HD double eps_of(double t) {   // MATLAB eps(t): spacing of the doubles at |t|
  const double a = fabs(t);
  if (a < 2.2250738585072014e-308) return 4.9406564584124654e-324;
  const int e = (int)((dbl_bits(a) >> 52) & 0x7ff) - 1023;
  return bits_dbl((int64_t)(e - 52 + 1023) << 52);
}

// Event location after an accepted step (EventZeroFcn, src/matlab/radau.m:1494-1618; the C++ port
// src/cpp/radaute_impl.hpp:2030-2185 is incomplete): E1 = event values at the start of the step, (T.t, T.y) its end.
// For every event whose value changed sign in the requested direction the zero is located on the collocation polynomial
// by the regula falsi / pegasus iteration of the reference; the earliest one is recorded.  Returns true when that
// event is terminal: the trajectory then ends at (te, ye).
template <int NY, int S, int SM>
JB_HD bool locate_events(Traj<NY, SM>& T, const RadauParams& P, const DenseOut& out, const double* E1v, double t1) {
  const double t2 = T.t;
  double tol = 65536.0 * fmax(EPS, fmax(eps_of(t2), eps_of(t1)));
  tol = fmin(tol, fabs(t2 - t1));
  bool found = false, stop = false;
  double tbest = 0.0, ybest[NY];
  int ibest = 0;
  for (int k = 0; k < P.ev_K; k++) {
    double E2 = -P.ev_b[k];
#pragma unroll
    for (int i = 0; i < NY; i++) E2 = fma(P.ev_A[k][i], T.y[i], E2);
    double t1N = t1, t2N = t2, E1 = E1v[k];
    const double p12 = (E2 - E1) / (t2N - t1N);
    if (!(E1 * E2 < 0 && p12 * P.ev_slope[k] >= 0)) continue;
    double tNew = t2N, ENew = E2, yNew[NY];
#pragma unroll
    for (int i = 0; i < NY; i++) yNew[i] = T.y[i];
    for (int iter = 1; iter < 50; iter++) {
      const bool tRel = fabs(t1N - t2N) * tol < fmax(fabs(t1N), fabs(t2N));
      const bool tAbs = fabs(t1N - t2N) < tol;
      if ((fabs(ENew) < tol && tRel && tAbs) || fabs(ENew) == 0.0) break;
      if (fabs(E1) < 200.0 * tol || fabs(E2) < 200.0 * tol) tNew = 0.5 * (t1N + t2N);
      else tNew = tNew + (-ENew / (E2 - E1) * (t2N - t1N));
      ntrprad<NY, S, SM>(T, tNew, yNew);
      ENew = -P.ev_b[k];
#pragma unroll
      for (int i = 0; i < NY; i++) ENew = fma(P.ev_A[k][i], yNew[i], ENew);
      if (ENew * E2 < 0) {
        t1N = t2N;
        E1 = E2;
      } else {
        E1 = E1 * E2 / (E2 + ENew);
      }
      t2N = tNew;
      E2 = ENew;
    }
    if (!found || (t1 < t2 ? tNew < tbest : tNew > tbest)) {
      found = true;
      tbest = tNew;
      ibest = k + 1;
      stop = P.ev_stop[k] != 0;
#pragma unroll
      for (int i = 0; i < NY; i++) ybest[i] = yNew[i];
    }
  }
  if (!found) return false;
  const int n = *out.nev;
  if (n < out.ecap) {
    out.te[n] = tbest;
    out.ie[n] = ibest;
#pragma unroll
    for (int i = 0; i < NY; i++) out.ye[n * NY + i] = ybest[i];
  }
  *out.nev = n + 1;
  if (stop) {
    T.t = tbest;
#pragma unroll
    for (int i = 0; i < NY; i++) T.y[i] = ybest[i];
  }
  return stop;
}

template <int NY, int S, int SM>
JB_HD void dense_output(Traj<NY, SM>& T, const DenseOut& out) {
  using Tab = RadauTab<S>;
  if (out.mode != 3) {   // OutFlag 1 / 2: the computed points (+ refine - 1 interpolated ones before each)
    for (int r = 1; r < out.refine; r++) {
      const double tint = T.t + T.h_old * ((double)r / (double)out.refine) - T.h_old;
      if (T.nout < out.cap) {
        out.tout[T.nout] = tint;
        ntrprad<NY, S, SM>(T, tint, out.yout + (long long)T.nout * NY);
      }
      T.nout += 1;
    }
    if (T.nout < out.cap) {
      out.tout[T.nout] = T.t;
#pragma unroll
      for (int i = 0; i < NY; i++) out.yout[(long long)T.nout * NY + i] = T.y[i];
    }
    T.nout += 1;
    return;
  }
  while (T.nout < out.ntspan) {
    const double tn = out.tspan[T.nout];
    if (T.PosNeg * (T.t - tn) < 0) break;
    if (T.t == tn) {
      out.tout[T.nout] = T.t;
#pragma unroll
      for (int i = 0; i < NY; i++) out.yout[T.nout * NY + i] = T.y[i];
      T.nout += 1;
      break;
    }
    const double s = (tn - T.t) / T.h_old;  // h_old = the step just taken
#pragma unroll
    for (int i = 0; i < NY; i++) {
      double yi = (s - (Tab::C(0) - 1.0)) * T.cont[(S - 1) * NY + i];
#pragma unroll
      for (int q = 1; q < S; q++) yi = (s - (Tab::C(q) - 1.0)) * (yi + T.cont[(S - 1 - q) * NY + i]);
      out.yout[T.nout * NY + i] = yi + T.y[i];
    }
    out.tout[T.nout] = tn;
    T.nout += 1;
  }
}

// radau.m:1267-1269: the last recorded point is overwritten with the final (t, y)
template <int NY, int SM>
JB_HD void dense_finish(const Traj<NY, SM>& T, const DenseOut& out) {
  if (out.mode != 3 && T.nout > out.cap) return;   // the last computed point was not stored
  if (T.nout >= 1) {
    out.tout[T.nout - 1] = T.t;
#pragma unroll
    for (int i = 0; i < NY; i++) out.yout[(T.nout - 1) * NY + i] = T.y[i];
  }
}

// ------------------------------------------------------------------ forward sensitivities
// S = d y / d y0 (or d y / d any ND parameters of the initial state), carried through every accepted
// step (the reference carries TensorDual numbers through RadauTeD::solve,
// /root/reference/src/cpp/radauted_impl.hpp:620-1645, with evalDynsDual / evalJacDual,
// src/cpp/janus_ode_common.hpp:2075-2158,2302-2416).  The dual part of the converged stage equations is
// LINEAR:   Zs_i = h sum_j a_ij J(y + z_j) (S + Zs_j),   y_{n+1}' = S + Zs_s,
// and is solved here by the same simplified-Newton iteration as the state -- same transformed system,
// same real + complex LU factors (the dual part of the iteration matrix only changes the convergence rate, not
// the fixed point, so no third derivatives of H are needed), with the Jacobians J(y + z_j) evaluated by
// hyper-dual numbers at the CONVERGED stage points.  The real part of the integration is untouched: states,
// step sizes and counters are bit-identical to a solve without sensitivities.
// Parameter directions (RadauParams::sens_par): the column s_k = d y / d par_k obeys the same linear stage system with
// the source term g_k(Y_j) = d f / d par_k added (ham_dpar: one more hyper-dual variable), s_k(t0) = 0 -- what the
// reference obtains by appending the parameter to the state (tests/cpp/ode_solver_test.cpp:630-640).
template <int NY>
struct Sens {
  static constexpr int ND = NY;
  double S[NY * NY];   // S[i * ND + d] = d y_i / d (direction d)
  double Sp[JB_SENS_MAXPAR * NY];   // Sp[k * NY + i] = d y_i / d ham_par[sens_par[k]]
  int iters;           // sensitivity Newton iterations spent (all directions, all steps)
  int unconverged;     // (direction, step) pairs whose iteration hit its cap or produced a non-finite increment: S is then
                       // not to be trusted -- reported to the caller instead of silently propagated
};

template <class Ham, int NY, int S, int SM, bool PAR>
JB_HD void sens_step_t(Traj<NY, SM>& T, const RadauParams& P, const double* z, const double* rscal, Sens<NY>& sn) {
  using Tab = RadauTab<S>;
  constexpr int ND = Sens<NY>::ND;
  constexpr int si = (S - 1) / 2;
  const int Nit = nit_of(S, P.MaxNbrNewton) + 10;
  const double tol = 0.1 * P.FNewt[si];
  const double rh = 1.0 / T.h;
  double Jq[S * NY * NY];
#pragma unroll
  for (int q = 0; q < S; q++) {
    double yq[NY];
#pragma unroll
    for (int i = 0; i < NY; i++) yq[i] = T.y[i] + z[q * NY + i];
    ham_jac<Ham>(yq, P.ham_par, P.sign, &Jq[q * NY * NY]);
  }
  T.work += (long long)S * Ham::FLOPS_J;
  for (int d = 0; d < ND + (PAR ? P.sens_npar : 0); d++) {
    double Sd[NY], smax = 1.0;
    double gq[PAR ? S * NY : 1];   // source term of a parameter direction at the stage points (unused for a state direction)
    if (!PAR || d < ND) {
#pragma unroll
      for (int i = 0; i < NY; i++) Sd[i] = sn.S[i * ND + d];
    } else if constexpr (PAR) {
#pragma unroll
      for (int i = 0; i < NY; i++) Sd[i] = sn.Sp[(d - ND) * NY + i];
#pragma unroll
      for (int q = 0; q < S; q++) {
        double yq[NY];
#pragma unroll
        for (int i = 0; i < NY; i++) yq[i] = T.y[i] + z[q * NY + i];
        ham_dpar<Ham>(yq, P.ham_par, P.sign, P.sens_par[d - ND], &gq[q * NY]);
      }
      T.work += (long long)S * Ham::FLOPS_J;
    }
#pragma unroll
    for (int i = 0; i < NY; i++) smax = fmax(smax, fabs(Sd[i]));
    const double rs = 1.0 / smax;
    double Zs[S * NY], Ws[S * NY];
#pragma unroll
    for (int k = 0; k < S * NY; k++) Zs[k] = Ws[k] = 0.0;
    double nrm_old = 0.0;
    for (int it = 0; it < Nit; it++) {
      double zz[S * NY];
      {  // F_q = J_q (S_d + Zs_q);  zz = TI F
        double Fq[S * NY];
#pragma unroll
        for (int q = 0; q < S; q++)
#pragma unroll
          for (int i = 0; i < NY; i++) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < NY; j++) acc = fma(Jq[q * NY * NY + i * NY + j], Sd[j] + Zs[q * NY + j], acc);
            Fq[q * NY + i] = (!PAR || d < ND) ? acc : acc + gq[PAR ? q * NY + i : 0];
          }
#pragma unroll
        for (int j = 0; j < S; j++)
#pragma unroll
          for (int i = 0; i < NY; i++) {
            double acc = Tab::TI(j, 0) * Fq[i];
#pragma unroll
            for (int q = 1; q < S; q++) acc = fma(Tab::TI(j, q), Fq[q * NY + i], acc);
            zz[j * NY + i] = acc;
          }
      }
      {  // Solvrad with the factors of this step
        const double valp0 = Tab::ValP(0) * rh;
#pragma unroll
        for (int i = 0; i < NY; i++) zz[i] = fma(-valp0, Ws[i], zz[i]);
        rlu_solve<NY>(T.L0, &zz[0]);
#pragma unroll
        for (int q = 1; q <= (S - 1) / 2; q++) {
          const int q2 = 2 * q - 1, q3 = 2 * q;
          const double v2 = Tab::ValP(q2) * rh, v3 = Tab::ValP(q3) * rh;
#pragma unroll
          for (int i = 0; i < NY; i++) {
            const double a = fma(v3, Ws[q3 * NY + i], fma(-v2, Ws[q2 * NY + i], zz[q2 * NY + i]));
            const double b = fma(-v2, Ws[q3 * NY + i], fma(-v3, Ws[q2 * NY + i], zz[q3 * NY + i]));
            zz[q2 * NY + i] = a;
            zz[q3 * NY + i] = b;
          }
          constexpr int NS = LcPlace<NY, SM>::NS;
          if constexpr (HiRoll<NY, S>::on) hi_solve<NY, SM>(T, q - 1, &zz[q2 * NY], &zz[q3 * NY]);
          else if ((q - 1) < NS) zlu_solve(T.LcS[(q - 1) < NS ? (q - 1) : 0], &zz[q2 * NY], &zz[q3 * NY]);
          else zlu_solve<NY>(T.Lc[(q - 1 - NS) > 0 ? (q - 1 - NS) : 0], &zz[q2 * NY], &zz[q3 * NY]);
        }
      }
      // Ws += dW;  Zs = T Ws;  size of the increment in Z space
      double nrm = 0.0;
#pragma unroll
      for (int k = 0; k < S * NY; k++) Ws[k] = Ws[k] + zz[k];
#pragma unroll
      for (int j = 0; j < S; j++)
#pragma unroll
        for (int i = 0; i < NY; i++) {
          double acc = Tab::T(j, 0) * Ws[i], dz = Tab::T(j, 0) * zz[i];
#pragma unroll
          for (int q = 1; q < S; q++) {
            acc = fma(Tab::T(j, q), Ws[q * NY + i], acc);
            dz = fma(Tab::T(j, q), zz[q * NY + i], dz);
          }
          Zs[j * NY + i] = acc;
          const double v = dz * rscal[i] * rs;
          nrm = fma(v, v, nrm);
        }
      nrm = ::sqrt(nrm) * P.inv_sqrt_sn[si];
      sn.iters += 1;
      T.work += W_newt(S, NY, 2 * NY * NY);
      if (bad(nrm)) {
        sn.unconverged += 1;
        break;
      }
      if (!(nrm > tol)) break;
      // the iteration is linear: increments contract by a fixed operator, so theta / (1 - theta) * |increment| bounds what
      // is left -- the state iteration's own stopping rule (radau.m:974-976, 1014), ten times tighter
      if (it >= 1) {
        const double theta = nrm / nrm_old;
        if (theta < 0.99 && theta / (1.0 - theta) * nrm <= tol) break;
      }
      nrm_old = nrm;
      if (it == Nit - 1) sn.unconverged += 1;
    }
    if (!PAR || d < ND) {
#pragma unroll
      for (int i = 0; i < NY; i++) sn.S[i * ND + d] = Sd[i] + Zs[(S - 1) * NY + i];
    } else {
#pragma unroll
      for (int i = 0; i < NY; i++) sn.Sp[(d - ND) * NY + i] = Sd[i] + Zs[(S - 1) * NY + i];
    }
  }
}

// the state directions alone keep the lean instantiation (no parameter source term in its inner loop)
template <class Ham, int NY, int S, int SM>
JB_HD void sens_step(Traj<NY, SM>& T, const RadauParams& P, const double* z, const double* rscal, Sens<NY>& sn) {
  if (P.sens_npar > 0) sens_step_t<Ham, NY, S, SM, true>(T, P, z, rscal, sn);
  else sens_step_t<Ham, NY, S, SM, false>(T, P, z, rscal, sn);
}

// factorisation of the rolled high-order steppers: real block in registers, complex blocks in a rolled loop into the
// compact run-time-indexed store (lu.cuh ZLUC); same arithmetic as decomrc
template <int NY, int S, int SM>
JB_HD void decomrc_hi(Traj<NY, SM>& T, const double* Jac) {
  using Tab = RadauTab<S>;
  const double rh0 = 1.0 / T.h;
  const double valp0 = Tab::ValP(0) * rh0;
#pragma unroll
  for (int i = 0; i < NY; i++)
#pragma unroll
    for (int j = 0; j < NY; j++) T.L0.a[i * NY + j] = (i == j) ? (valp0 - Jac[i * NY + j]) : -Jac[i * NY + j];
  rlu_factor<NY>(T.L0);
#pragma unroll 1
  for (int q = 1; q <= (S - 1) / 2; q++) {
    const double al = Tab::ValP(2 * q - 1) * rh0, be = Tab::ValP(2 * q) * rh0;
    ZLU<NY> Z;
#pragma unroll
    for (int i = 0; i < NY; i++)
#pragma unroll
      for (int j = 0; j < NY; j++) {
        Z.re[i * NY + j] = (i == j) ? (al - Jac[i * NY + j]) : -Jac[i * NY + j];
        Z.im[i * NY + j] = (i == j) ? be : 0.0;
      }
    zlu_factor<NY>(Z);
    hi_store<NY, SM>(T, q - 1, Z);
  }
}

// ------------------------------------------------------------------ rolled high-order attempt (5 and 7 stages)
// Same operations in the same order as attempt_step below -- bit-identical results -- but organised for the
// instruction cache: the loops over the complex blocks (factorisation, Solvrad) and over the output stage of the
// T / TI transforms are ROLLED, the transformed stage vectors z and w live in local memory (dynamically indexed;
// the unrolled version spilled them anyway: 2.1 kB of spill stores per step at 7 stages) and all complex factors in
// shared memory (lu.cuh: ZLUC).  The 7-stage stepper shrinks from ~16,000 to ~5,000 SASS instructions.
template <class Ham, int NY, int S, int SM>
JB_HD bool attempt_step_hi(Traj<NY, SM>& T, const RadauParams& P, double* Jac, bool jac_fresh, const DenseOut* out,
                           Sens<NY>* sens) {
  using Tab = RadauTab<S>;
  constexpr int si = (S - 1) / 2, NQ = (S - 1) / 2;
  const int Nit = nit_of(S, P.MaxNbrNewton);
  double zl[S * NY], wl[S * NY];  // [stage][component], local memory

  if (T.NeedNewLU) {  // radau.m:827-844
    if (!jac_fresh) ham_jac<Ham>(T.yJ, P.ham_par, P.sign, Jac);
    decomrc_hi<NY, S, SM>(T, Jac);
    T.st[ST_DEC] += 1;
    T.work += W_lu(S, NY);
    T.NeedNewLU = false;
  }
  const bool Variab = (P.MaxNbrStg - P.MinNbrStg) != 0;
  if (Variab && T.Keep) {  // radau.m:846-853
    T.Keep = false;
    T.ChangeNbr += 1;
    if (T.ChangeNbr >= 10 && S < P.MaxNbrStg) T.NeedNewJac = false;
  }
  if (0.1 * fabs(T.h) <= fabs(T.t) * EPS) {  // radau.m:855
    T.res = 1;
    return true;
  }
  double rscal[NY];
#pragma unroll
  for (int i = 0; i < NY; i++) {
    const double sc = fma(P.RelTol1[si][i], fabs(T.y[i]), P.AbsTol1[si][i]);  // radau.m:858-862
    rscal[i] = 1.0 / sc;
  }
  if (T.First || P.Start_Newt || T.ChangeFlag) {  // radau.m:872-891
#pragma unroll 1
    for (int k = 0; k < S * NY; k++) {
      zl[k] = 0.0;
      wl[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < S * NY; k++) T.cont[k] = 0.0;
  } else {
    const double hquot = T.h / T.h_old;
#pragma unroll 1
    for (int q = 0; q < S; q++) {
      const double cq = Tab::C(q) * hquot;
#pragma unroll
      for (int i = 0; i < NY; i++) {
        double zq = (cq - Tab::C(0) + 1.0) * T.cont[(S - 1) * NY + i];
#pragma unroll
        for (int q1 = 2; q1 <= S; q1++) zq = (zq + T.cont[(S - q1) * NY + i]) * (cq - Tab::C(q1 - 1) + 1.0);
        zl[q * NY + i] = zq;
      }
    }
    double zr[S * NY];
#pragma unroll
    for (int k = 0; k < S * NY; k++) zr[k] = zl[k];
#pragma unroll 1
    for (int j = 0; j < S; j++)
#pragma unroll
      for (int i = 0; i < NY; i++) {
        double acc = Tab::TI(j, 0) * zr[i];
#pragma unroll
        for (int q = 1; q < S; q++) acc = fma(Tab::TI(j, q), zr[q * NY + i], acc);
        wl[j * NY + i] = acc;
      }
  }
  // ------- simplified Newton (radau.m:897-1018)
  const double FNewt = P.FNewt[si];
  T.FacConv = det_pow(fmax(T.FacConv, EPS), 0.8);
  T.Theta = fabs(P.Thet);
  T.Newt = 0;
  const double rh = 1.0 / T.h;
  double NewNrm = 0.0, OldNrm = 0.0, thqold = 0.0;
  bool newton_failed = false;
  for (;;) {
    T.Reject = false;  // radau.m:913 (quirk Q9)
    T.Newt += 1;
    if (T.Newt > Nit) {  // radau.m:915-925
      T.UnExpStepRej = true;
      T.st[ST_REJ] += 1;
      T.st[ST_NREJ] += 1;
      T.h = T.h * 0.5;
      T.Reject = true;
      T.Last = false;
      T.NeedNewLU = true;
      newton_failed = true;
      break;
    }
    {  // stage right-hand sides f(y + z_q), then z = TI f  (radau.m:932-948)
      double fq[S * NY];
      bool nanf = false;
#pragma unroll
      for (int q = 0; q < S; q++) {
        double yq[NY];
#pragma unroll
        for (int i = 0; i < NY; i++) yq[i] = T.y[i] + zl[q * NY + i];
        ham_dyns<Ham>(yq, P.ham_par, P.sign, &fq[q * NY]);
#pragma unroll
        for (int i = 0; i < NY; i++) nanf = nanf || (fq[q * NY + i] != fq[q * NY + i]);
      }
      if (nanf) {
        T.res = 2;
        return true;
      }
      T.st[ST_FCN] += S;
#pragma unroll 1
      for (int j = 0; j < S; j++)
#pragma unroll
        for (int i = 0; i < NY; i++) {
          double acc = Tab::TI(j, 0) * fq[i];
#pragma unroll
          for (int q = 1; q < S; q++) acc = fma(Tab::TI(j, q), fq[q * NY + i], acc);
          zl[j * NY + i] = acc;
        }
    }
    {  // Solvrad (radaute_impl.hpp:1721-1868)
      const double valp0 = Tab::ValP(0) * rh;
      double z0[NY];
#pragma unroll
      for (int i = 0; i < NY; i++) z0[i] = fma(-valp0, wl[i], zl[i]);
      rlu_solve<NY>(T.L0, z0);
#pragma unroll
      for (int i = 0; i < NY; i++) zl[i] = z0[i];
#pragma unroll 1
      for (int q = 1; q <= NQ; q++) {
        const int q2 = 2 * q - 1, q3 = 2 * q;
        const double v2 = Tab::ValP(q2) * rh, v3 = Tab::ValP(q3) * rh;
        double a[NY], b[NY];
#pragma unroll
        for (int i = 0; i < NY; i++) {
          const double w2 = wl[q2 * NY + i], w3 = wl[q3 * NY + i];
          a[i] = fma(v3, w3, fma(-v2, w2, zl[q2 * NY + i]));
          b[i] = fma(-v2, w3, fma(-v3, w2, zl[q3 * NY + i]));
        }
        hi_solve<NY, SM>(T, q - 1, a, b);
#pragma unroll
        for (int i = 0; i < NY; i++) {
          zl[q2 * NY + i] = a[i];
          zl[q3 * NY + i] = b[i];
        }
      }
    }
    T.st[ST_SOL] += 1;
    T.work += W_newt(S, NY, Ham::FLOPS_F);
    NewNrm = 0.0;  // radau.m:955-959
#pragma unroll 1
    for (int q = 0; q < S; q++) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < NY; i++) {
        const double v = zl[q * NY + i] * rscal[i];
        s = fma(v, v, s);
      }
      NewNrm += ::sqrt(s);
    }
    NewNrm = NewNrm * P.inv_sqrt_sn[si];
    if (bad(NewNrm)) {
      T.res = 2;
      return true;
    }
    if (T.Newt > 1 && T.Newt < Nit) {  // radau.m:962-996
      const double thq = NewNrm / OldNrm;
      if (T.Newt == 2) T.Theta = thq;
      else T.Theta = ::sqrt(thq * thqold);
      thqold = thq;
      if (T.Theta < 0.99) {
        T.FacConv = T.Theta / (1.0 - T.Theta);
        const double dyth = T.FacConv * NewNrm * ipow(T.Theta, Nit - 1 - T.Newt) / FNewt;
        if (dyth >= 1.0) {
          const double qNewt = fmax(1.0e-4, fmin(20.0, dyth));
          const double hhfac = 0.8 * det_pow(qNewt, -1.0 / (4.0 + Nit - 1 - T.Newt));
          T.h = hhfac * T.h;
          T.Reject = true;
          T.Last = false;
          T.UnExpNewtRej = hhfac <= 0.5;
          T.st[ST_NREJ] += 1;
          T.st[ST_REJ] += 1;
          T.NeedNewLU = true;
          newton_failed = true;
          break;
        }
      } else {
        T.h = T.h * 0.5;
        T.Reject = true;
        T.Last = false;
        T.UnExpStepRej = true;
        T.st[ST_REJ] += 1;
        T.NeedNewLU = true;
        newton_failed = true;
        break;
      }
    }
    OldNrm = fmax(NewNrm, EPS);  // radau.m:1001-1012
    {
      double wr[S * NY];
#pragma unroll
      for (int k = 0; k < S * NY; k++) {
        wr[k] = wl[k] + zl[k];
        wl[k] = wr[k];
      }
#pragma unroll 1
      for (int j = 0; j < S; j++)
#pragma unroll
        for (int i = 0; i < NY; i++) {
          double acc = Tab::T(j, 0) * wr[i];
#pragma unroll
          for (int q = 1; q < S; q++) acc = fma(Tab::T(j, q), wr[q * NY + i], acc);
          zl[j * NY + i] = acc;
        }
    }
    if (T.FacConv * NewNrm > FNewt) continue;  // radau.m:1014
    break;
  }
  if (newton_failed) return false;  // radau.m:1020-1022

  // ------- Estrad (radaute_impl.hpp:1902-1986)
  T.work += W_tail(S, NY, Ham::FLOPS_F);
  double temp[NY], ev[NY];
#pragma unroll
  for (int i = 0; i < NY; i++) temp[i] = zl[i] * (Tab::Dd(0) * rh);
#pragma unroll 1
  for (int q = 1; q < S; q++) {
    const double d = Tab::Dd(q) * rh;
#pragma unroll
    for (int i = 0; i < NY; i++) temp[i] = fma(zl[q * NY + i], d, temp[i]);
  }
#pragma unroll
  for (int i = 0; i < NY; i++) ev[i] = T.f0[i] + temp[i];
  rlu_solve<NY>(T.L0, ev);
  double err;
  {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NY; i++) {
      const double v = ev[i] * rscal[i];
      s = fma(v, v, s);
    }
    err = fmax(::sqrt(s) * P.inv_sqrt_n, 1e-10);
  }
  if (err >= 1.0 && (T.First || T.Reject)) {
    double ya[NY];
#pragma unroll
    for (int i = 0; i < NY; i++) ya[i] = T.y[i] + err;  // sic: scalar err (radaute_impl.hpp:1951)
    ham_dyns<Ham>(ya, P.ham_par, P.sign, ev);
    T.st[ST_FCN] += 1;
    T.work += Ham::FLOPS_F + 2LL * NY * NY + 3LL * NY;
#pragma unroll
    for (int i = 0; i < NY; i++) ev[i] = ev[i] + temp[i];
    rlu_solve<NY>(T.L0, ev);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NY; i++) {
      const double v = ev[i] * rscal[i];
      s = fma(v, v, s);
    }
    err = fmax(::sqrt(s) * P.inv_sqrt_n, 1e-10);
  }
  if (bad(err)) {
    T.res = 2;
    return true;
  }
  // ------- new step size (radau.m:1038-1050)
  const double fac = fmin(P.Safe, (2.0 * Nit + 1.0) / (2.0 * Nit + T.Newt));
  double quot = fmax(P.FacR, fmin(P.FacL, det_pow(err, 1.0 / S + 1.0) / fac));
  double hnew = T.h / quot;
  if (err < 1.0) {  // ------- accepted (radau.m:1054-1232)
    T.First = false;
    T.st[ST_ACC] += 1;
    const bool gus_upd = (P.Gustafsson & 1) && !T.ChangeFlag;   // quirk Q8: see attempt_step
    const bool gus_lim = (P.Gustafsson & 2) ? true : gus_upd;
    if (gus_lim) {
      if (T.st[ST_ACC] > 1) {
        const double ratio = err * err / T.erracc;
        double root;
        if (S == 7) root = ::sqrt(::sqrt(::sqrt(ratio)));
        else root = det_pow(ratio, 1.0 / (S + 1.0));
        double facgus = (T.hacc / T.h) * root / P.Safe;
        facgus = fmax(P.FacR, fmin(P.FacL, facgus));
        quot = fmax(quot, facgus);
        hnew = T.h / quot;
      }
    }
    if (gus_upd) {
      T.hacc = T.h;
      T.erracc = fmax(1.0e-2, err);
    }
    if (sens) sens_step<Ham, NY, S, SM>(T, P, zl, rscal, *sens);
    double E1v[4];
    const double t_ev1 = T.t;
    if (out && P.ev_K > 0) {   // event values at the start of the step (the reference keeps them from its previous call)
      for (int k = 0; k < P.ev_K; k++) {
        double e = -P.ev_b[k];
#pragma unroll
        for (int i = 0; i < NY; i++) e = fma(P.ev_A[k][i], T.y[i], e);
        E1v[k] = e;
      }
    }
    T.h_old = T.h;
    T.t = T.t + T.h;
    double z[S * NY];
#pragma unroll
    for (int k = 0; k < S * NY; k++) z[k] = zl[k];
#pragma unroll
    for (int i = 0; i < NY; i++) T.y[i] = T.y[i] + z[(S - 1) * NY + i];
    // collocation polynomial (radau.m:1093-1107)
#pragma unroll
    for (int i = 0; i < NY; i++) T.cont[(S - 1) * NY + i] = z[i] * (1.0 / Tab::C(0));
#pragma unroll
    for (int q = 1; q <= S - 1; q++) {
      const double Fact = 1.0 / (Tab::C(S - q - 1) - Tab::C(S - q));
#pragma unroll
      for (int i = 0; i < NY; i++) T.cont[(q - 1) * NY + i] = (z[(S - q - 1) * NY + i] - z[(S - q) * NY + i]) * Fact;
    }
#pragma unroll
    for (int jj = 2; jj <= S; jj++)
#pragma unroll
      for (int k = S; k >= jj; k--) {
        const double Fact = (S - k == 0) ? 1.0 / (-Tab::C(jj - 1)) : 1.0 / (Tab::C(S - k - 1) - Tab::C(S - k + jj - 1));
#pragma unroll
        for (int i = 0; i < NY; i++) T.cont[(k - 1) * NY + i] = (T.cont[(k - 1) * NY + i] - T.cont[(k - 2) * NY + i]) * Fact;
      }
    if (out && P.ev_K > 0 && locate_events<NY, S, SM>(T, P, *out, E1v, t_ev1)) return true;   // terminal event (radau.m:1110-1121)
    if (out) dense_output<NY, S, SM>(T, *out);  // radau.m:1150-1168
    T.NeedNewJac = true;
    if (T.Last) {  // radau.m:1193-1197
      T.h = T.hopt;
      T.st[ST_REJ] += 1;
      return true;
    }
    ham_dyns<Ham>(T.y, P.ham_par, P.sign, T.f0);
    T.st[ST_FCN] += 1;
    {
      bool nanf = false;
#pragma unroll
      for (int i = 0; i < NY; i++) nanf = nanf || (T.f0[i] != T.f0[i]);
      if (nanf) {
        T.res = 2;
        return true;
      }
    }
    hnew = T.PosNeg * fmin(fabs(hnew), fabs(T.hmaxn));
    T.hopt = T.PosNeg * fmin(fabs(T.h), fabs(hnew));
    if (T.Reject) hnew = T.PosNeg * fmin(fabs(hnew), fabs(T.h));
    T.Reject = false;
    if ((T.t + hnew / P.Quot1 - T.tfinal) * T.PosNeg >= 0) {
      T.h = T.tfinal - T.t;
      T.Last = true;
    } else {
      const double qt = hnew / T.h;
      if (T.Theta <= P.Thet && qt >= P.Quot1 && qt <= P.Quot2) {
        T.Keep = true;
        T.NeedNewJac = false;
        T.NeedNewLU = false;
        return false;
      }
      T.h = hnew;
    }
    T.NeedNewLU = true;
    if (T.Theta <= P.Thet) T.NeedNewJac = false;
  } else {  // ------- rejected (radau.m:1234-1253)
    T.Reject = true;
    T.Last = false;
    if (T.First) T.h = T.h / 10.0;
    else T.h = hnew;
    if (T.st[ST_ACC] >= 1) T.st[ST_REJ] += 1;
    T.NeedNewLU = true;
  }
  return false;
}

template <class Ham, int NY, int S, int SM>
JB_HD bool attempt_step(Traj<NY, SM>& T, const RadauParams& P, double* Jac, bool jac_fresh, const DenseOut* out = nullptr,
                        Sens<NY>* sens = nullptr) {
  if constexpr (HiRoll<NY, S>::on) return attempt_step_hi<Ham, NY, S, SM>(T, P, Jac, jac_fresh, out, sens);
  using Tab = RadauTab<S>;
  constexpr int si = (S - 1) / 2;
  const int Nit = nit_of(S, P.MaxNbrNewton);
  double z[S * NY], w[S * NY];  // [stage][component]

  if (T.NeedNewLU) {  // radau.m:827-844
    if (!jac_fresh) ham_jac<Ham>(T.yJ, P.ham_par, P.sign, Jac);  // the Jacobian kept from an earlier pass
    decomrc<NY, S, SM>(T, Jac);
    T.st[ST_DEC] += 1;
    T.work += W_lu(S, NY);
    T.NeedNewLU = false;
  }
  const bool Variab = (P.MaxNbrStg - P.MinNbrStg) != 0;
  if (Variab && T.Keep) {  // radau.m:846-853
    T.Keep = false;
    T.ChangeNbr += 1;
    if (T.ChangeNbr >= 10 && S < P.MaxNbrStg) T.NeedNewJac = false;
  }
  if (0.1 * fabs(T.h) <= fabs(T.t) * EPS) {  // radau.m:855
    T.res = 1;
    return true;
  }
  double rscal[NY];
#pragma unroll
  for (int i = 0; i < NY; i++) {
    const double sc = fma(P.RelTol1[si][i], fabs(T.y[i]), P.AbsTol1[si][i]);  // radau.m:858-862
    rscal[i] = 1.0 / sc;
  }
  if (T.First || P.Start_Newt || T.ChangeFlag) {  // radau.m:872-891
#pragma unroll
    for (int k = 0; k < S * NY; k++) {
      z[k] = 0.0;
      w[k] = 0.0;
      T.cont[k] = 0.0;
    }
  } else {
    const double hquot = T.h / T.h_old;
#pragma unroll
    for (int q = 0; q < S; q++) {
      const double cq = Tab::C(q) * hquot;
#pragma unroll
      for (int i = 0; i < NY; i++) {
        double zq = (cq - Tab::C(0) + 1.0) * T.cont[(S - 1) * NY + i];
#pragma unroll
        for (int q1 = 2; q1 <= S; q1++) zq = (zq + T.cont[(S - q1) * NY + i]) * (cq - Tab::C(q1 - 1) + 1.0);
        z[q * NY + i] = zq;
      }
    }
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
      for (int i = 0; i < NY; i++) {
        double acc = Tab::TI(j, 0) * z[i];
#pragma unroll
        for (int q = 1; q < S; q++) acc = fma(Tab::TI(j, q), z[q * NY + i], acc);
        w[j * NY + i] = acc;
      }
  }
  // ------- simplified Newton (radau.m:897-1018)
  const double FNewt = P.FNewt[si];
  T.FacConv = det_pow(fmax(T.FacConv, EPS), 0.8);
  T.Theta = fabs(P.Thet);
  T.Newt = 0;
  const double rh = 1.0 / T.h;
  double NewNrm = 0.0, OldNrm = 0.0, thqold = 0.0;
  bool newton_failed = false;
  for (;;) {
    T.Reject = false;  // radau.m:913 (quirk Q9)
    T.Newt += 1;
    if (T.Newt > Nit) {  // radau.m:915-925
      T.UnExpStepRej = true;
      T.st[ST_REJ] += 1;
      T.st[ST_NREJ] += 1;
      T.h = T.h * 0.5;
      T.Reject = true;
      T.Last = false;
      T.NeedNewLU = true;
      newton_failed = true;
      break;
    }
    // stage right-hand sides f(y + z_q), then z = TI f  (radau.m:932-948)
    {
      double fq[S * NY];
      bool nanf = false;
#pragma unroll
      for (int q = 0; q < S; q++) {
        double yq[NY];
#pragma unroll
        for (int i = 0; i < NY; i++) yq[i] = T.y[i] + z[q * NY + i];
        ham_dyns<Ham>(yq, P.ham_par, P.sign, &fq[q * NY]);
#pragma unroll
        for (int i = 0; i < NY; i++) nanf = nanf || (fq[q * NY + i] != fq[q * NY + i]);
      }
      if (nanf) {
        T.res = 2;
        return true;
      }
      T.st[ST_FCN] += S;
#pragma unroll
      for (int j = 0; j < S; j++)
#pragma unroll
        for (int i = 0; i < NY; i++) {
          double acc = Tab::TI(j, 0) * fq[i];
#pragma unroll
          for (int q = 1; q < S; q++) acc = fma(Tab::TI(j, q), fq[q * NY + i], acc);
          z[j * NY + i] = acc;
        }
    }
    // Solvrad (radaute_impl.hpp:1721-1868)
    {
      const double valp0 = Tab::ValP(0) * rh;
#pragma unroll
      for (int i = 0; i < NY; i++) z[i] = fma(-valp0, w[i], z[i]);
      rlu_solve<NY>(T.L0, &z[0]);
#pragma unroll
      for (int q = 1; q <= (S - 1) / 2; q++) {
        const int q2 = 2 * q - 1, q3 = 2 * q;
        const double v2 = Tab::ValP(q2) * rh, v3 = Tab::ValP(q3) * rh;
#pragma unroll
        for (int i = 0; i < NY; i++) {
          const double a = fma(v3, w[q3 * NY + i], fma(-v2, w[q2 * NY + i], z[q2 * NY + i]));
          const double b = fma(-v2, w[q3 * NY + i], fma(-v3, w[q2 * NY + i], z[q3 * NY + i]));
          z[q2 * NY + i] = a;
          z[q3 * NY + i] = b;
        }
        constexpr int NS = LcPlace<NY, SM>::NS;
        if ((q - 1) < NS) zlu_solve(T.LcS[(q - 1) < NS ? (q - 1) : 0], &z[q2 * NY], &z[q3 * NY]);
        else zlu_solve<NY>(T.Lc[(q - 1 - NS) > 0 ? (q - 1 - NS) : 0], &z[q2 * NY], &z[q3 * NY]);
      }
    }
    T.st[ST_SOL] += 1;
    T.work += W_newt(S, NY, Ham::FLOPS_F);
    NewNrm = 0.0;  // radau.m:955-959
#pragma unroll
    for (int q = 0; q < S; q++) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < NY; i++) {
        const double v = z[q * NY + i] * rscal[i];
        s = fma(v, v, s);
      }
      NewNrm += ::sqrt(s);
    }
    NewNrm = NewNrm * P.inv_sqrt_sn[si];
    if (bad(NewNrm)) {
      T.res = 2;
      return true;
    }
    if (T.Newt > 1 && T.Newt < Nit) {  // radau.m:962-996
      const double thq = NewNrm / OldNrm;
      if (T.Newt == 2) T.Theta = thq;
      else T.Theta = ::sqrt(thq * thqold);
      thqold = thq;
      if (T.Theta < 0.99) {
        T.FacConv = T.Theta / (1.0 - T.Theta);
        const double dyth = T.FacConv * NewNrm * ipow(T.Theta, Nit - 1 - T.Newt) / FNewt;
        if (dyth >= 1.0) {
          const double qNewt = fmax(1.0e-4, fmin(20.0, dyth));
          const double hhfac = 0.8 * det_pow(qNewt, -1.0 / (4.0 + Nit - 1 - T.Newt));
          T.h = hhfac * T.h;
          T.Reject = true;
          T.Last = false;
          T.UnExpNewtRej = hhfac <= 0.5;
          T.st[ST_NREJ] += 1;
          T.st[ST_REJ] += 1;
          T.NeedNewLU = true;
          newton_failed = true;
          break;
        }
      } else {
        T.h = T.h * 0.5;
        T.Reject = true;
        T.Last = false;
        T.UnExpStepRej = true;
        T.st[ST_REJ] += 1;
        T.NeedNewLU = true;
        newton_failed = true;
        break;
      }
    }
    OldNrm = fmax(NewNrm, EPS);  // radau.m:1001-1012
#pragma unroll
    for (int k = 0; k < S * NY; k++) w[k] = w[k] + z[k];
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
      for (int i = 0; i < NY; i++) {
        double acc = Tab::T(j, 0) * w[i];
#pragma unroll
        for (int q = 1; q < S; q++) acc = fma(Tab::T(j, q), w[q * NY + i], acc);
        z[j * NY + i] = acc;
      }
    if (T.FacConv * NewNrm > FNewt) continue;  // radau.m:1014
    break;
  }
  if (newton_failed) return false;  // radau.m:1020-1022

  // ------- Estrad (radaute_impl.hpp:1902-1986)
  T.work += W_tail(S, NY, Ham::FLOPS_F);
  double temp[NY], ev[NY];
#pragma unroll
  for (int i = 0; i < NY; i++) {
    double s = z[i] * (Tab::Dd(0) * rh);
#pragma unroll
    for (int q = 1; q < S; q++) s = fma(z[q * NY + i], Tab::Dd(q) * rh, s);
    temp[i] = s;
    ev[i] = T.f0[i] + s;
  }
  rlu_solve<NY>(T.L0, ev);
  double err;
  {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NY; i++) {
      const double v = ev[i] * rscal[i];
      s = fma(v, v, s);
    }
    err = fmax(::sqrt(s) * P.inv_sqrt_n, 1e-10);
  }
  if (err >= 1.0 && (T.First || T.Reject)) {
    double ya[NY];
#pragma unroll
    for (int i = 0; i < NY; i++) ya[i] = T.y[i] + err;  // sic: scalar err (radaute_impl.hpp:1951)
    ham_dyns<Ham>(ya, P.ham_par, P.sign, ev);
    T.st[ST_FCN] += 1;
    T.work += Ham::FLOPS_F + 2LL * NY * NY + 3LL * NY;
#pragma unroll
    for (int i = 0; i < NY; i++) ev[i] = ev[i] + temp[i];
    rlu_solve<NY>(T.L0, ev);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NY; i++) {
      const double v = ev[i] * rscal[i];
      s = fma(v, v, s);
    }
    err = fmax(::sqrt(s) * P.inv_sqrt_n, 1e-10);
  }
  if (bad(err)) {
    T.res = 2;
    return true;
  }
  // ------- new step size (radau.m:1038-1050)
  const double fac = fmin(P.Safe, (2.0 * Nit + 1.0) / (2.0 * Nit + T.Newt));
  double quot = fmax(P.FacR, fmin(P.FacL, det_pow(err, 1.0 / S + 1.0) / fac));
  double hnew = T.h / quot;
  if (err < 1.0) {  // ------- accepted (radau.m:1054-1232)
    T.First = false;
    T.st[ST_ACC] += 1;
    // P.Gustafsson: bit 0 = predictive controller on; bit 1 = RadauTe's gate: the limiter is applied on
    // every accepted step with AccptNbr > 1, also when the order just changed (radaute_impl.hpp:1092-1100: the inner
    // mask m1_12_1_1_1 lacks `& ~ChangeFlag`), while hacc / erracc are only refreshed under `Gustafsson & ~ChangeFlag`
    // (:1101-1102).  Bit 1 clear = radau.m:1066-1076, which gates both.
    const bool gus_upd = (P.Gustafsson & 1) && !T.ChangeFlag;
    const bool gus_lim = (P.Gustafsson & 2) ? true : gus_upd;
    if (gus_lim) {
      if (T.st[ST_ACC] > 1) {
        const double ratio = err * err / T.erracc;
        double root;
        if (S == 1) root = ::sqrt(ratio);
        else if (S == 3) root = ::sqrt(::sqrt(ratio));
        else if (S == 7) root = ::sqrt(::sqrt(::sqrt(ratio)));
        else root = det_pow(ratio, 1.0 / (S + 1.0));
        double facgus = (T.hacc / T.h) * root / P.Safe;
        facgus = fmax(P.FacR, fmin(P.FacL, facgus));
        quot = fmax(quot, facgus);
        hnew = T.h / quot;
      }
    }
    if (gus_upd) {
      T.hacc = T.h;
      T.erracc = fmax(1.0e-2, err);
    }
    if (sens) sens_step<Ham, NY, S, SM>(T, P, z, rscal, *sens);   // before y moves: needs y_n, z, the LU of this step
    double E1v[4];
    const double t_ev1 = T.t;
    if (out && P.ev_K > 0) {   // event values at the start of the step (the reference keeps them from its previous call)
      for (int k = 0; k < P.ev_K; k++) {
        double e = -P.ev_b[k];
#pragma unroll
        for (int i = 0; i < NY; i++) e = fma(P.ev_A[k][i], T.y[i], e);
        E1v[k] = e;
      }
    }
    T.h_old = T.h;
    T.t = T.t + T.h;
#pragma unroll
    for (int i = 0; i < NY; i++) T.y[i] = T.y[i] + z[(S - 1) * NY + i];
    // collocation polynomial (radau.m:1093-1107)
#pragma unroll
    for (int i = 0; i < NY; i++) T.cont[(S - 1) * NY + i] = z[i] * (1.0 / Tab::C(0));
#pragma unroll
    for (int q = 1; q <= S - 1; q++) {
      const double Fact = 1.0 / (Tab::C(S - q - 1) - Tab::C(S - q));
#pragma unroll
      for (int i = 0; i < NY; i++) T.cont[(q - 1) * NY + i] = (z[(S - q - 1) * NY + i] - z[(S - q) * NY + i]) * Fact;
    }
#pragma unroll
    for (int jj = 2; jj <= S; jj++)
#pragma unroll
      for (int k = S; k >= jj; k--) {
        const double Fact = (S - k == 0) ? 1.0 / (-Tab::C(jj - 1)) : 1.0 / (Tab::C(S - k - 1) - Tab::C(S - k + jj - 1));
#pragma unroll
        for (int i = 0; i < NY; i++) T.cont[(k - 1) * NY + i] = (T.cont[(k - 1) * NY + i] - T.cont[(k - 2) * NY + i]) * Fact;
      }
    if (out && P.ev_K > 0 && locate_events<NY, S, SM>(T, P, *out, E1v, t_ev1)) return true;   // terminal event (radau.m:1110-1121)
    if (out) dense_output<NY, S, SM>(T, *out);  // radau.m:1150-1168
    T.NeedNewJac = true;
    if (T.Last) {  // radau.m:1193-1197
      T.h = T.hopt;
      T.st[ST_REJ] += 1;
      return true;
    }
    ham_dyns<Ham>(T.y, P.ham_par, P.sign, T.f0);
    T.st[ST_FCN] += 1;
    {
      bool nanf = false;
#pragma unroll
      for (int i = 0; i < NY; i++) nanf = nanf || (T.f0[i] != T.f0[i]);
      if (nanf) {
        T.res = 2;
        return true;
      }
    }
    hnew = T.PosNeg * fmin(fabs(hnew), fabs(T.hmaxn));
    T.hopt = T.PosNeg * fmin(fabs(T.h), fabs(hnew));
    if (T.Reject) hnew = T.PosNeg * fmin(fabs(hnew), fabs(T.h));
    T.Reject = false;
    if ((T.t + hnew / P.Quot1 - T.tfinal) * T.PosNeg >= 0) {
      T.h = T.tfinal - T.t;
      T.Last = true;
    } else {
      const double qt = hnew / T.h;
      if (T.Theta <= P.Thet && qt >= P.Quot1 && qt <= P.Quot2) {
        T.Keep = true;
        T.NeedNewJac = false;
        T.NeedNewLU = false;
        return false;
      }
      T.h = hnew;
    }
    T.NeedNewLU = true;
    if (T.Theta <= P.Thet) T.NeedNewJac = false;
  } else {  // ------- rejected (radau.m:1234-1253)
    T.Reject = true;
    T.Last = false;
    if (T.First) T.h = T.h / 10.0;
    else T.h = hnew;
    if (T.st[ST_ACC] >= 1) T.st[ST_REJ] += 1;
    T.NeedNewLU = true;
  }
  return false;
}

// head of a pass: Jacobian refresh + order selection (radau.m:733-824); returns false if the
// trajectory must stop (NaN in the Jacobian)
template <class Ham, int NY, int SM, bool VAR>
JB_HD bool pass_head(Traj<NY, SM>& T, const RadauParams& P, double* Jac, bool& jac_fresh) {
  T.st[ST_STEP] += 1;
  T.FacConv = det_pow(fmax(T.FacConv, EPS), 0.8);
  if (T.NeedNewJac) {
    ham_jac<Ham>(T.y, P.ham_par, P.sign, Jac);
    jac_fresh = true;
#pragma unroll
    for (int i = 0; i < NY; i++) T.yJ[i] = T.y[i];
    bool nanf = false;
#pragma unroll
    for (int k = 0; k < NY * NY; k++) nanf = nanf || (Jac[k] != Jac[k]);
    if (nanf) {
      T.res = 2;
      return false;
    }
    T.st[ST_JAC] += 1;
    T.work += Ham::FLOPS_J;
    T.NeedNewJac = false;
    T.NeedNewLU = true;
  }
  const bool Variab = (P.MaxNbrStg - P.MinNbrStg) != 0;
  if (VAR && Variab && !T.Keep) {
    T.ChangeNbr += 1;
    int NbrStgNew = T.NbrStg;
    const double hquot = T.h / T.h_old;
    T.Thetat = fmin(10.0, fmax(T.Theta, T.Thetat * 0.5));
    if (T.Newt > 1 && T.Thetat <= P.Vitu && hquot < P.hhou && hquot > P.hhod) NbrStgNew = imin(P.MaxNbrStg, T.NbrStg + 2);
    if (T.Thetat >= P.Vitd || T.UnExpStepRej) NbrStgNew = imax(P.MinNbrStg, T.NbrStg - 2);
    if (T.ChangeNbr >= 1 && T.UnExpNewtRej) NbrStgNew = imax(P.MinNbrStg, T.NbrStg - 2);
    if (T.ChangeNbr <= 10) NbrStgNew = imin(T.NbrStg, NbrStgNew);
    T.ChangeFlag = T.NbrStg != NbrStgNew;
    T.UnExpNewtRej = false;
    T.UnExpStepRej = false;
    if (T.ChangeFlag) {
      T.NbrStg = NbrStgNew;
      T.ChangeNbr = 1;
      T.NeedNewLU = true;
    }
  }
  return true;
}

// whole integration of one trajectory.  VAR = false: fixed order S = SM (host guarantees
// MinNbrStg == MaxNbrStg == NbrStg == SM); VAR = true: orders 1..SM selected at run time.
template <class Ham, int NY, int SM, bool VAR>
JB_HD bool one_pass(Traj<NY, SM>& T, const RadauParams& P, const DenseOut* out = nullptr, Sens<NY>* sens = nullptr) {
  double Jac[NY * NY];
  bool jf = false;
  if (!pass_head<Ham, NY, SM, VAR>(T, P, Jac, jf)) return true;
  if (!VAR) return attempt_step<Ham, NY, SM, SM>(T, P, Jac, jf, out, sens);
  switch (T.NbrStg) {
    case 1: return attempt_step<Ham, NY, 1, SM>(T, P, Jac, jf, out, sens);
    case 3: return attempt_step<Ham, NY, (SM >= 3 ? 3 : 1), SM>(T, P, Jac, jf, out, sens);
    case 5: return attempt_step<Ham, NY, (SM >= 5 ? 5 : 1), SM>(T, P, Jac, jf, out, sens);
    default: return attempt_step<Ham, NY, (SM >= 7 ? 7 : 1), SM>(T, P, Jac, jf, out, sens);
  }
}

JB_HD bool traj_active(double PosNeg, double t, double tfinal, int steps, long long MaxNbrStep) {
  return steps <= MaxNbrStep && PosNeg * t < PosNeg * tfinal;
}

template <class Ham, int NY, int SM, bool VAR>
JB_HD void integrate(Traj<NY, SM>& T, const RadauParams& P, const DenseOut* out = nullptr, Sens<NY>* sens = nullptr) {
  if (T.res != 0) return;
  while (traj_active(T.PosNeg, T.t, T.tfinal, T.st[ST_STEP], P.MaxNbrStep)) {
    if (one_pass<Ham, NY, SM, VAR>(T, P, out, sens)) return;
  }
  if (T.res == 0 && T.PosNeg * T.t < T.PosNeg * T.tfinal) T.res = 3;  // MaxNbrStep exceeded
}


// ------------------------------------------------------------------ parked trajectories
// Variable-order solves regroup trajectories by stage count between "epochs" so that every warp
// runs ONE attempt_step<S> instantiation (order-homogeneous warps; the active-set compaction of
// the north star).  Between epochs a trajectory lives in HBM as a Parked record.  The LU factors
// are NOT stored: they are a deterministic function of (Jac, h) and are recomputed on resume
// (same code, same bits); z / w are rebuilt from cont by the starting-value extrapolation.
enum : unsigned {
  PF_LAST = 1u, PF_FIRST = 2u, PF_REJECT = 4u, PF_KEEP = 8u, PF_CHANGE = 16u, PF_UNEXP_STEP = 32u,
  PF_UNEXP_NEWT = 64u, PF_NEED_JAC = 128u, PF_NEED_LU = 256u, PF_HEAD_DONE = 512u, PF_FINISHED = 1024u
};

template <int NY>
struct alignas(16) Parked {
  double y[NY], f0[NY], yJ[NY], cont[7 * NY];
  double t, t0, tfinal, h, h_old, hopt, hacc, erracc, FacConv, Theta, Thetat;
  long long work;
  int st[8];
  int NbrStg, Newt, ChangeNbr, res;
  unsigned flags;
  int nout;
};

template <int NY, int SM>
JB_HD void park(const Traj<NY, SM>& T, double t0, bool head_done, bool finished, Parked<NY>& p) {
#pragma unroll
  for (int i = 0; i < NY; i++) { p.y[i] = T.y[i]; p.f0[i] = T.f0[i]; p.yJ[i] = T.yJ[i]; }
#pragma unroll
  for (int i = 0; i < SM * NY; i++) p.cont[i] = T.cont[i];
  p.t = T.t; p.t0 = t0; p.tfinal = T.tfinal; p.h = T.h; p.h_old = T.h_old; p.hopt = T.hopt;
  p.hacc = T.hacc; p.erracc = T.erracc; p.FacConv = T.FacConv; p.Theta = T.Theta; p.Thetat = T.Thetat;
  p.work = T.work;
#pragma unroll
  for (int i = 0; i < 8; i++) p.st[i] = T.st[i];
  p.NbrStg = T.NbrStg; p.Newt = T.Newt; p.ChangeNbr = T.ChangeNbr; p.res = T.res;
  p.flags = (T.Last ? PF_LAST : 0u) | (T.First ? PF_FIRST : 0u) | (T.Reject ? PF_REJECT : 0u) |
            (T.Keep ? PF_KEEP : 0u) | (T.ChangeFlag ? PF_CHANGE : 0u) | (T.UnExpStepRej ? PF_UNEXP_STEP : 0u) |
            (T.UnExpNewtRej ? PF_UNEXP_NEWT : 0u) | (T.NeedNewJac ? PF_NEED_JAC : 0u) |
            (T.NeedNewLU ? PF_NEED_LU : 0u) | (head_done ? PF_HEAD_DONE : 0u) | (finished ? PF_FINISHED : 0u);
  p.nout = T.nout;
}

template <int NY, int SM>
JB_HD void unpark(const Parked<NY>& p, Traj<NY, SM>& T, double& t0, bool& head_done) {
#pragma unroll
  for (int i = 0; i < NY; i++) { T.y[i] = p.y[i]; T.f0[i] = p.f0[i]; T.yJ[i] = p.yJ[i]; }
#pragma unroll
  for (int i = 0; i < SM * NY; i++) T.cont[i] = p.cont[i];
  T.t = p.t; t0 = p.t0; T.tfinal = p.tfinal; T.h = p.h; T.h_old = p.h_old; T.hopt = p.hopt;
  T.hacc = p.hacc; T.erracc = p.erracc; T.FacConv = p.FacConv; T.Theta = p.Theta; T.Thetat = p.Thetat;
  T.PosNeg = (p.tfinal - p.t0) >= 0 ? 1.0 : -1.0;
  T.hmaxn = fabs(p.tfinal - p.t0);
  T.work = p.work;
#pragma unroll
  for (int i = 0; i < 8; i++) T.st[i] = p.st[i];
  T.NbrStg = p.NbrStg; T.Newt = p.Newt; T.ChangeNbr = p.ChangeNbr; T.res = p.res;
  T.Last = p.flags & PF_LAST; T.First = p.flags & PF_FIRST; T.Reject = p.flags & PF_REJECT;
  T.Keep = p.flags & PF_KEEP; T.ChangeFlag = p.flags & PF_CHANGE; T.UnExpStepRej = p.flags & PF_UNEXP_STEP;
  T.UnExpNewtRej = p.flags & PF_UNEXP_NEWT; T.NeedNewJac = p.flags & PF_NEED_JAC; T.NeedNewLU = p.flags & PF_NEED_LU;
  head_done = p.flags & PF_HEAD_DONE;
  T.nout = p.nout;
}

// Run one parked trajectory at stage count S for at most max_passes attempted steps.
// Returns true when the trajectory is finished; otherwise it must be re-parked (epoch budget
// used up, or its order changed -> head_done tells the next owner to skip pass_head).
// stop_on_accept (incremental step() API): also stop, with *stepped = true, right after the first
// accepted step.
// rebuild the factors a trajectory was parked with (first thing after unpark)
template <class Ham, int NY, int S>
JB_HD void epoch_begin(Traj<NY, S>& T, const RadauParams& P) {
  if (!T.NeedNewLU) {
    double Jac[NY * NY];
    ham_jac<Ham>(T.yJ, P.ham_par, P.sign, Jac);
    if constexpr (HiRoll<NY, S>::on) decomrc_hi<NY, S, S>(T, Jac);
    else decomrc<NY, S, S>(T, Jac);
  }
}

// one pass of run_epoch's loop: 0 = go on, 1 = trajectory finished, 2 = must be re-parked (its order changed ->
// head_done, or stop_on_accept hit -> *stepped)
template <class Ham, int NY, int S, bool VAR = true>
JB_HD int epoch_pass(Traj<NY, S>& T, const RadauParams& P, bool& head_done, int acc0, const DenseOut* out = nullptr,
                     bool stop_on_accept = false, bool* stepped = nullptr, Sens<NY>* sens = nullptr) {
  if (!(T.res == 0 && traj_active(T.PosNeg, T.t, T.tfinal, T.st[ST_STEP], P.MaxNbrStep))) {
    if (T.res == 0 && T.PosNeg * T.t < T.PosNeg * T.tfinal) T.res = 3;
    return 1;
  }
  double Jac[NY * NY];
  bool jf = false;
  if (!head_done) {
    if (!pass_head<Ham, NY, S, VAR>(T, P, Jac, jf)) return 1;
  }
  if (T.NbrStg != S) {
    head_done = true;
    return 2;
  }
  head_done = false;
  if (attempt_step<Ham, NY, S, S>(T, P, Jac, jf, out, sens)) return 1;
  if (stop_on_accept && T.st[ST_ACC] != acc0) {
    *stepped = true;
    return 2;
  }
  return 0;
}

// Run one parked trajectory at stage count S for at most max_passes attempted steps.
// Returns true when the trajectory is finished; otherwise it must be re-parked (epoch budget
// used up, or its order changed -> head_done tells the next owner to skip pass_head).
// stop_on_accept (incremental step() API): also stop, with *stepped = true, right after the first
// accepted step.
template <class Ham, int NY, int S, bool VAR = true>
JB_HD bool run_epoch(Traj<NY, S>& T, const RadauParams& P, bool& head_done, int max_passes,
                     const DenseOut* out = nullptr, bool stop_on_accept = false, bool* stepped = nullptr,
                     Sens<NY>* sens = nullptr) {
  epoch_begin<Ham, NY, S>(T, P);
  const int acc0 = T.st[ST_ACC];
  for (int passes = 0;;) {
    const int r = epoch_pass<Ham, NY, S, VAR>(T, P, head_done, acc0, out, stop_on_accept, stepped, sens);
    if (r) return r == 1;
    if (++passes >= max_passes) return false;
  }
}

}  // namespace jb
