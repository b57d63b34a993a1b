// Radau IIA stepper for Hamiltonians whose stage evaluation is a batched GEMM (the MLP of config 4):
// the per-trajectory state machine of radau_core.cuh cut at every dynamics / Jacobian evaluation
// into PHASES, so that the evaluations of all trajectories run together as tile GEMMs
// (mlp_kernels.cuh) while everything else -- LU, Newton update, error estimate, controller, order
// selection -- stays one thread per trajectory on structure-of-arrays state in HBM.
//
//   reference: /root/reference/src/cpp/radaute_impl.hpp:537-1345 (solve), :1502-1661 (DecomRC),
//   :1721-1897 (Solvrad), :1902-1986 (Estrad); scalar truth src/matlab/radau.m:712-1256.
//
// Every floating-point operation is the one radau_core.cuh::attempt_step / pass_head performs, in
// the same order (ARITHMETIC CONTRACT), so a trajectory integrated here is bit-identical to the
// same trajectory integrated by the thread-resident path and by the mirror oracle.  The file is
// __host__ __device__ throughout: tests/hostsim runs the same phases on the CPU.
//
// Phase protocol of one pass (= one attempted step of every active trajectory):
//   ph_head      loop guard, step counter, FacConv; asks for a Jacobian when NeedNewJac
//   [mlp_hess_kernel over the Jacobian list]
//   ph_prep      Jacobian bookkeeping, order selection, LU, starting values, first newton_begin
//   repeat:  [mlp_grad_kernel over the (trajectory, stage) rows]   ph_newton
//   [mlp_grad_kernel at y + err]  ph_estrad_b         (first-step retry of Estrad only)
//   [mlp_grad_kernel at the new y, saving activations]  ph_accept_b   (after an accepted step)
#pragma once
#include "mlp_common.cuh"
#include "radau_core.cuh"

namespace jb {
namespace mlp {

constexpr int MN = 16;   // Ny of the MLP Hamiltonian

enum Phase : int {
  PH_IDLE = 0,      // between passes
  PH_WANT_JAC,      // ph_head -> Jacobian kernel -> ph_prep
  PH_PREP,          // ph_head -> ph_prep (Jacobian still valid)
  PH_NEWTON,        // stage rows requested; ph_newton consumes F
  PH_ERR_EVAL,      // Estrad retry: f(y + err) requested into ev
  PH_F0_EVAL,       // accepted: f(t_new, y_new) requested into f0
  PH_FINISHED
};

struct TrajS {   // per-trajectory scalars (array of structs)
  double t, t0, tfinal, h, h_old, hopt, hacc, erracc, FacConv, Theta, Thetat, PosNeg, hmaxn;
  double hnew, err, NewNrm, OldNrm, thqold;
  long long work;
  int st[8];
  int NbrStg, Newt, ChangeNbr, res;
  unsigned flags;   // PF_* of radau_core.cuh
  int phase;
  int nout;         // output points written so far (output at the tspan points, OutFlag 3)
  int acc_mark;     // incremental protocol: accepted-step count when the current step() call began
};

// vector state: structure of arrays, element e of trajectory i at arr[e * M + i] (one thread per
// trajectory: coalesced).  The matrices are arrays of structures -- Jac[i][256] and one LU block per
// trajectory -- because they are produced / factored by 16 lanes per trajectory, one row per lane.
struct State {
  long long M;
  int SM;          // stage capacity (MaxNbrStg)
  TrajS* sc;
  double *y, *f0, *cont, *z, *w, *F, *rscal, *temp, *ev, *errs;
  double* Jac;     // [i][16][16]
  double* LU;      // [i][lud]: L0a[256] L0rd[16] { re[256] im[256] rdre[16] rdim[16] } x (SM-1)/2
  int* LUpiv;      // [i][lui]: L0piv[16] { piv[16] } x (SM-1)/2
  int lud, lui;
  // output at the interior points of tspan (radau.m:1150-1168, OutFlag 3); tout == nullptr: final state only
  int ntspan;              // columns of tspan (>= 2)
  const double* tspan;     // [M][ntspan]
  double* tout;            // [M][ntspan]
  double* yout;            // [M][ntspan][16]
  JB_HD double* L0a(long long i) const { return LU + (size_t)i * lud; }
  JB_HD double* L0rd(long long i) const { return LU + (size_t)i * lud + 256; }
  JB_HD int* L0piv(long long i) const { return LUpiv + (size_t)i * lui; }
  JB_HD double* Lre(long long i, int q) const { return LU + (size_t)i * lud + 272 + q * 544; }
  JB_HD double* Lim(long long i, int q) const { return Lre(i, q) + 256; }
  JB_HD double* Lrdre(long long i, int q) const { return Lre(i, q) + 512; }
  JB_HD double* Lrdim(long long i, int q) const { return Lre(i, q) + 528; }
  JB_HD int* Lpiv(long long i, int q) const { return LUpiv + (size_t)i * lui + 16 + q * 16; }
};
JB_HD int lu_doubles(int SM) { return 272 + ((SM - 1) / 2) * 544; }
JB_HD int lu_ints(int SM) { return 16 + ((SM - 1) / 2) * 16; }
constexpr unsigned PF_LU_PENDING = 2048u;   // ph_prep(defer_lu): factorisation left to the 16-lane LU kernel

#define JB_FLAG(s, f) (((s).flags & (f)) != 0u)
JB_HD void set_flag(TrajS& s, unsigned f, bool v) { s.flags = v ? (s.flags | f) : (s.flags & ~f); }

// ------------------------------------------------------------------ small dense LU on SoA columns
// Factorisation and forward substitution: same operations, same order as lu.cuh (rows exchanged for
// real instead of by predicated selects: identical values).  Back substitution: column-oriented
// (ARITHMETIC CONTRACT of the Ny = 16 path, mirrored by oracle/c rlu_solve / zlu_solve for n = 16).
JB_HD void rlu_factor_g(double* a, double* rd, int* piv, long long M) {
  constexpr int n = MN;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(a[(size_t)(k * n + k) * M]);
    for (int i = k + 1; i < n; i++) {
      const double v = fabs(a[(size_t)(i * n + k) * M]);
      if (v > best) { best = v; p = i; }
    }
    piv[(size_t)k * M] = p;
    if (p != k)
      for (int j = 0; j < n; j++) {
        const double x = a[(size_t)(k * n + j) * M];
        a[(size_t)(k * n + j) * M] = a[(size_t)(p * n + j) * M];
        a[(size_t)(p * n + j) * M] = x;
      }
    const double r = 1.0 / a[(size_t)(k * n + k) * M];
    rd[(size_t)k * M] = r;
    for (int i = k + 1; i < n; i++) {
      const double l = a[(size_t)(i * n + k) * M] * r;
      a[(size_t)(i * n + k) * M] = l;
      for (int j = k + 1; j < n; j++)
        a[(size_t)(i * n + j) * M] = fma(-l, a[(size_t)(k * n + j) * M], a[(size_t)(i * n + j) * M]);
    }
  }
}

JB_HD void rlu_solve_g(const double* a, const double* rd, const int* piv, long long LS /*matrix element stride*/,
                       long long M /*vector element stride*/, double* b) {
  constexpr int n = MN;
  for (int k = 0; k < n; k++) {
    const int p = piv[(size_t)k * LS];
    if (p != k) {
      const double x = b[(size_t)k * M];
      b[(size_t)k * M] = b[(size_t)p * M];
      b[(size_t)p * M] = x;
    }
  }
  for (int i = 1; i < n; i++) {
    double s = b[(size_t)i * M];
    for (int j = 0; j < i; j++) s = fma(-a[(size_t)(i * n + j) * LS], b[(size_t)j * M], s);
    b[(size_t)i * M] = s;
  }
  // column-oriented back substitution (the order the 16-lane solve of lu16.cuh produces; also the
  // order of reference-BLAS dtrsm): unknown j is finished, then every row above it is updated
  for (int j = n - 1; j >= 0; j--) {
    const double xj = b[(size_t)j * M] * rd[(size_t)j * LS];
    b[(size_t)j * M] = xj;
    for (int i = 0; i < j; i++) b[(size_t)i * M] = fma(-a[(size_t)(i * n + j) * LS], xj, b[(size_t)i * M]);
  }
}

JB_HD void zlu_factor_g(double* re, double* im, double* rdre, double* rdim, int* piv, long long M) {
  constexpr int n = MN;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(re[(size_t)(k * n + k) * M]) + fabs(im[(size_t)(k * n + k) * M]);
    for (int i = k + 1; i < n; i++) {
      const double v = fabs(re[(size_t)(i * n + k) * M]) + fabs(im[(size_t)(i * n + k) * M]);
      if (v > best) { best = v; p = i; }
    }
    piv[(size_t)k * M] = p;
    if (p != k)
      for (int j = 0; j < n; j++) {
        double x = re[(size_t)(k * n + j) * M];
        re[(size_t)(k * n + j) * M] = re[(size_t)(p * n + j) * M];
        re[(size_t)(p * n + j) * M] = x;
        x = im[(size_t)(k * n + j) * M];
        im[(size_t)(k * n + j) * M] = im[(size_t)(p * n + j) * M];
        im[(size_t)(p * n + j) * M] = x;
      }
    const double c = re[(size_t)(k * n + k) * M], d = im[(size_t)(k * n + k) * M];
    const double inv = 1.0 / fma(d, d, c * c);
    const double rr = c * inv, ri = -d * inv;
    rdre[(size_t)k * M] = rr;
    rdim[(size_t)k * M] = ri;
    for (int i = k + 1; i < n; i++) {
      const double xr = re[(size_t)(i * n + k) * M], xi = im[(size_t)(i * n + k) * M];
      const double lr = fma(-xi, ri, xr * rr), li = fma(xi, rr, xr * ri);
      re[(size_t)(i * n + k) * M] = lr;
      im[(size_t)(i * n + k) * M] = li;
      for (int j = k + 1; j < n; j++) {
        const double ur = re[(size_t)(k * n + j) * M], ui = im[(size_t)(k * n + j) * M];
        re[(size_t)(i * n + j) * M] = fma(li, ui, fma(-lr, ur, re[(size_t)(i * n + j) * M]));
        im[(size_t)(i * n + j) * M] = fma(-li, ur, fma(-lr, ui, im[(size_t)(i * n + j) * M]));
      }
    }
  }
}

JB_HD void zlu_solve_g(const double* re, const double* im, const double* rdre, const double* rdim, const int* piv,
                       long long LS, long long M, double* br, double* bi) {
  constexpr int n = MN;
  for (int k = 0; k < n; k++) {
    const int p = piv[(size_t)k * LS];
    if (p != k) {
      double x = br[(size_t)k * M];
      br[(size_t)k * M] = br[(size_t)p * M];
      br[(size_t)p * M] = x;
      x = bi[(size_t)k * M];
      bi[(size_t)k * M] = bi[(size_t)p * M];
      bi[(size_t)p * M] = x;
    }
  }
  for (int i = 1; i < n; i++) {
    double sr = br[(size_t)i * M], si = bi[(size_t)i * M];
    for (int j = 0; j < i; j++) {
      const double lr = re[(size_t)(i * n + j) * LS], li = im[(size_t)(i * n + j) * LS];
      const double bj = br[(size_t)j * M], cj = bi[(size_t)j * M];
      sr = fma(li, cj, fma(-lr, bj, sr));
      si = fma(-li, bj, fma(-lr, cj, si));
    }
    br[(size_t)i * M] = sr;
    bi[(size_t)i * M] = si;
  }
  for (int j = n - 1; j >= 0; j--) {   // column-oriented, see rlu_solve_g
    const double rr = rdre[(size_t)j * LS], ri = rdim[(size_t)j * LS];
    const double bj = br[(size_t)j * M], cj = bi[(size_t)j * M];
    const double xr = fma(-cj, ri, bj * rr), xi = fma(cj, rr, bj * ri);
    br[(size_t)j * M] = xr;
    bi[(size_t)j * M] = xi;
    for (int i = 0; i < j; i++) {
      const double ur = re[(size_t)(i * n + j) * LS], ui = im[(size_t)(i * n + j) * LS];
      const double sr = br[(size_t)i * M], si = bi[(size_t)i * M];
      br[(size_t)i * M] = fma(ui, xi, fma(-ur, xr, sr));
      bi[(size_t)i * M] = fma(-ui, xr, fma(-ur, xi, si));
    }
  }
}

// ------------------------------------------------------------------ phases
JB_HD void finish(TrajS& s) { s.phase = PH_FINISHED; }

// radau.m:631-696 / radaute_impl.hpp:406-455 up to the first dynamics evaluation
JB_HD void ph_init(const State& S, const RadauParams& P, long long i, const double* y0 /*[M][16]*/) {
  const double* tspan = S.tspan + (size_t)i * S.ntspan;
  const long long M = S.M;
  TrajS s;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) S.y[(size_t)e * M + i] = y0[i * MN + e];
  for (int e = 0; e < S.SM * MN; e++) S.cont[(size_t)e * M + i] = 0.0;
  for (int k = 0; k < 8; k++) s.st[k] = 0;
  const double t0 = tspan[0], tfinal = tspan[S.ntspan - 1];
  s.nout = 0;
  s.t = t0;
  s.t0 = t0;
  s.tfinal = tfinal;
  s.PosNeg = (tfinal - t0) >= 0 ? 1.0 : -1.0;
  s.NbrStg = P.NbrStg;
  s.flags = 0u;
  s.ChangeNbr = 0;
  s.Theta = s.Thetat = 0.0;
  s.hmaxn = fabs(tfinal - t0);
  double h = P.InitialStep;
  if (fabs(h) <= 10 * EPS) h = 1e-6;
  h = s.PosNeg * fmin(fabs(h), s.hmaxn);
  s.h_old = h;
  s.hopt = h;
  if ((s.t + h * 1.0001 - tfinal) * s.PosNeg >= 0) {
    h = tfinal - s.t;
    s.flags |= PF_LAST;
  }
  s.h = h;
  s.FacConv = 1.0;
  s.flags |= PF_FIRST | PF_NEED_JAC | PF_NEED_LU;
  s.st[ST_FCN] += 1;
  s.work = FLOPS_F;
  s.res = 0;
  s.Newt = 0;
  s.hacc = 0.0;
  s.erracc = 0.0;
  s.hnew = s.err = s.NewNrm = s.OldNrm = s.thqold = 0.0;
  s.acc_mark = 0;
  s.phase = PH_F0_EVAL;
  S.sc[i] = s;
}

JB_HD void ph_init_b(const State& S, long long i) {
  const long long M = S.M;
  TrajS& s = S.sc[i];
  bool nanf = false;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) {
    const double v = S.f0[(size_t)e * M + i];
    nanf = nanf || (v != v);
  }
  if (nanf) s.res = 2;
  s.phase = nanf ? PH_FINISHED : PH_IDLE;
}

// incremental protocol (RadauT.step(htry, ...), src/python/seulext.py:158-581): start of a step() call -- an optional
// trial step chosen by the caller (like radau.m:645-655; same rule as rearm_kernel of the thread-resident path)
JB_HD void ph_rearm(const State& S, long long i, const double* htry) {
  TrajS& s = S.sc[i];
  if (s.phase == PH_FINISHED) return;
  if (htry && htry[i] != s.h) {
    double h = s.PosNeg * fmin(fabs(htry[i]), fabs(s.tfinal - s.t0));
    s.flags |= PF_NEED_LU;
    if ((s.t + h * 1.0001 - s.tfinal) * s.PosNeg >= 0) {
      h = s.tfinal - s.t;
      s.flags |= PF_LAST;
    } else {
      s.flags &= ~PF_LAST;
    }
    s.h = h;
  }
  s.acc_mark = s.st[ST_ACC];
}

// run_epoch's loop guard + the first lines of pass_head (radau.m:712, :733-735)
JB_HD void ph_head(const State& S, const RadauParams& P, long long i, bool stop_on_accept = false) {
  TrajS& s = S.sc[i];
  if (s.phase == PH_FINISHED) return;
  if (stop_on_accept && s.st[ST_ACC] != s.acc_mark) return;   // incremental protocol: this call's step is done
  if (!(s.res == 0 && traj_active(s.PosNeg, s.t, s.tfinal, s.st[ST_STEP], P.MaxNbrStep))) {
    if (s.res == 0 && s.PosNeg * s.t < s.PosNeg * s.tfinal) s.res = 3;
    s.phase = PH_FINISHED;
    return;
  }
  s.st[ST_STEP] += 1;
  s.FacConv = det_pow(fmax(s.FacConv, EPS), 0.8);
  s.phase = JB_FLAG(s, PF_NEED_JAC) ? PH_WANT_JAC : PH_PREP;
}

template <int S_>
JB_HD void newton_begin(TrajS& s, const RadauParams& P) {
  const int Nit = nit_of(S_, P.MaxNbrNewton);
  set_flag(s, PF_REJECT, false);   // radau.m:913 (quirk Q9)
  s.Newt += 1;
  if (s.Newt > Nit) {              // radau.m:915-925
    s.flags |= PF_UNEXP_STEP;
    s.st[ST_REJ] += 1;
    s.st[ST_NREJ] += 1;
    s.h = s.h * 0.5;
    s.flags |= PF_REJECT | PF_NEED_LU;
    s.flags &= ~PF_LAST;
    s.phase = PH_IDLE;
    return;
  }
  s.phase = PH_NEWTON;
}

template <int S_>
JB_HD void decomrc_g(const State& S, const TrajS& s, long long i) {
  using Tab = RadauTab<S_>;
  const double rh = 1.0 / s.h;
  const double valp0 = Tab::ValP(0) * rh;
  const double* Jac = S.Jac + (size_t)i * MN * MN;
  double* a0 = S.L0a(i);
  for (int r = 0; r < MN; r++)
    for (int c = 0; c < MN; c++) {
      const double J = Jac[r * MN + c];
      a0[r * MN + c] = (r == c) ? (valp0 - J) : -J;
    }
  rlu_factor_g(a0, S.L0rd(i), S.L0piv(i), 1);
#pragma unroll
  for (int q = 1; q <= (S_ - 1) / 2; q++) {
    const double al = Tab::ValP(2 * q - 1) * rh, be = Tab::ValP(2 * q) * rh;
    double* re = S.Lre(i, q - 1);
    double* im = S.Lim(i, q - 1);
    for (int r = 0; r < MN; r++)
      for (int c = 0; c < MN; c++) {
        const double J = Jac[r * MN + c];
        re[r * MN + c] = (r == c) ? (al - J) : -J;
        im[r * MN + c] = (r == c) ? be : 0.0;
      }
    zlu_factor_g(re, im, S.Lrdre(i, q - 1), S.Lrdim(i, q - 1), S.Lpiv(i, q - 1), 1);
  }
}

// attempt_step up to the first stage evaluation (radau.m:827-911)
template <int S_>
JB_HD void attempt_begin(const State& S, const RadauParams& P, TrajS& s, long long i, bool defer_lu) {
  using Tab = RadauTab<S_>;
  constexpr int si = (S_ - 1) / 2;
  const long long M = S.M;
  if (JB_FLAG(s, PF_NEED_LU)) {
    if (defer_lu) s.flags |= PF_LU_PENDING;   // h and the stage count are final from here on
    else decomrc_g<S_>(S, s, i);
    s.st[ST_DEC] += 1;
    s.work += W_lu(S_, MN);
    s.flags &= ~PF_NEED_LU;
  }
  const bool Variab = (P.MaxNbrStg - P.MinNbrStg) != 0;
  if (Variab && JB_FLAG(s, PF_KEEP)) {
    s.flags &= ~PF_KEEP;
    s.ChangeNbr += 1;
    if (s.ChangeNbr >= 10 && S_ < P.MaxNbrStg) s.flags &= ~PF_NEED_JAC;
  }
  if (0.1 * fabs(s.h) <= fabs(s.t) * EPS) {
    s.res = 1;
    finish(s);
    return;
  }
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) {
    const double sc = fma(P.RelTol1[si][e], fabs(S.y[(size_t)e * M + i]), P.AbsTol1[si][e]);
    S.rscal[(size_t)e * M + i] = 1.0 / sc;
  }
  if (JB_FLAG(s, PF_FIRST) || P.Start_Newt || JB_FLAG(s, PF_CHANGE)) {
    for (int k = 0; k < S_ * MN; k++) {
      S.z[(size_t)k * M + i] = 0.0;
      S.w[(size_t)k * M + i] = 0.0;
      S.cont[(size_t)k * M + i] = 0.0;
    }
  } else {
    const double hquot = s.h / s.h_old;
#pragma unroll
    for (int q = 0; q < S_; q++) {
      const double cq = Tab::C(q) * hquot;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
      for (int e = 0; e < MN; e++) {
        double zq = (cq - Tab::C(0) + 1.0) * S.cont[(size_t)((S_ - 1) * MN + e) * M + i];
#pragma unroll
        for (int q1 = 2; q1 <= S_; q1++)
          zq = (zq + S.cont[(size_t)((S_ - q1) * MN + e) * M + i]) * (cq - Tab::C(q1 - 1) + 1.0);
        S.z[(size_t)(q * MN + e) * M + i] = zq;
      }
    }
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
    for (int e = 0; e < MN; e++) {
      double zz[S_];
#pragma unroll
      for (int q = 0; q < S_; q++) zz[q] = S.z[(size_t)(q * MN + e) * M + i];
#pragma unroll
      for (int j = 0; j < S_; j++) {
        double acc = Tab::TI(j, 0) * zz[0];
#pragma unroll
        for (int q = 1; q < S_; q++) acc = fma(Tab::TI(j, q), zz[q], acc);
        S.w[(size_t)(j * MN + e) * M + i] = acc;
      }
    }
  }
  s.FacConv = det_pow(fmax(s.FacConv, EPS), 0.8);
  s.Theta = fabs(P.Thet);
  s.Newt = 0;
  s.NewNrm = s.OldNrm = s.thqold = 0.0;
  newton_begin<S_>(s, P);
}

// the factorisation ph_prep(defer_lu = true) left pending, scalar form (the GPU runs lu16_kernel instead)
JB_HD void ph_lu(const State& S, long long i) {
  TrajS& s = S.sc[i];
  if (!JB_FLAG(s, PF_LU_PENDING)) return;
  s.flags &= ~PF_LU_PENDING;
  switch (s.NbrStg) {
    case 1: decomrc_g<1>(S, s, i); break;
    case 3: decomrc_g<3>(S, s, i); break;
    case 5: decomrc_g<5>(S, s, i); break;
    default: decomrc_g<7>(S, s, i); break;
  }
}

// the rest of pass_head (Jacobian bookkeeping, order selection radau.m:766-824), then attempt_begin
JB_HD void ph_prep(const State& S, const RadauParams& P, long long i, bool defer_lu = false) {
  const long long M = S.M;
  TrajS s = S.sc[i];
  if (s.phase != PH_WANT_JAC && s.phase != PH_PREP) return;
  if (s.phase == PH_WANT_JAC) {
    bool nanf = false;
    for (int k = 0; k < MN * MN; k++) {
      const double v = S.Jac[(size_t)i * MN * MN + k];
      nanf = nanf || (v != v);
    }
    if (nanf) {
      s.res = 2;
      finish(s);
      S.sc[i] = s;
      return;
    }
    s.st[ST_JAC] += 1;
    s.work += FLOPS_J;
    s.flags &= ~PF_NEED_JAC;
    s.flags |= PF_NEED_LU;
  }
  const bool Variab = (P.MaxNbrStg - P.MinNbrStg) != 0;
  if (Variab && !JB_FLAG(s, PF_KEEP)) {
    s.ChangeNbr += 1;
    int NbrStgNew = s.NbrStg;
    const double hquot = s.h / s.h_old;
    s.Thetat = fmin(10.0, fmax(s.Theta, s.Thetat * 0.5));
    if (s.Newt > 1 && s.Thetat <= P.Vitu && hquot < P.hhou && hquot > P.hhod) NbrStgNew = imin(P.MaxNbrStg, s.NbrStg + 2);
    if (s.Thetat >= P.Vitd || JB_FLAG(s, PF_UNEXP_STEP)) NbrStgNew = imax(P.MinNbrStg, s.NbrStg - 2);
    if (s.ChangeNbr >= 1 && JB_FLAG(s, PF_UNEXP_NEWT)) NbrStgNew = imax(P.MinNbrStg, s.NbrStg - 2);
    if (s.ChangeNbr <= 10) NbrStgNew = imin(s.NbrStg, NbrStgNew);
    set_flag(s, PF_CHANGE, s.NbrStg != NbrStgNew);
    s.flags &= ~(PF_UNEXP_NEWT | PF_UNEXP_STEP);
    if (JB_FLAG(s, PF_CHANGE)) {
      s.NbrStg = NbrStgNew;
      s.ChangeNbr = 1;
      s.flags |= PF_NEED_LU;
    }
  }
  switch (s.NbrStg) {
    case 1: attempt_begin<1>(S, P, s, i, defer_lu); break;
    case 3: attempt_begin<3>(S, P, s, i, defer_lu); break;
    case 5: attempt_begin<5>(S, P, s, i, defer_lu); break;
    default: attempt_begin<7>(S, P, s, i, defer_lu); break;
  }
  S.sc[i] = s;
}

// after an accepted step: every requested time in (t - h, t] evaluated on the collocation polynomial
// (ntrprad, radaute_impl.hpp:1413-1440; same loop as radau_core.cuh dense_output)
template <int S_>
JB_HD void dense_output_g(const State& S, TrajS& s, long long i) {
  using Tab = RadauTab<S_>;
  const long long M = S.M;
  const double* tspan = S.tspan + (size_t)i * S.ntspan;
  double* tout = S.tout + (size_t)i * S.ntspan;
  double* yout = S.yout + (size_t)i * S.ntspan * MN;
  while (s.nout < S.ntspan) {
    const double tn = tspan[s.nout];
    if (s.PosNeg * (s.t - tn) < 0) break;
    if (s.t == tn) {
      tout[s.nout] = s.t;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
      for (int e = 0; e < MN; e++) yout[s.nout * MN + e] = S.y[(size_t)e * M + i];
      s.nout += 1;
      break;
    }
    const double sg = (tn - s.t) / s.h_old;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
    for (int e = 0; e < MN; e++) {
      double yi = (sg - (Tab::C(0) - 1.0)) * S.cont[(size_t)((S_ - 1) * MN + e) * M + i];
#pragma unroll
      for (int q = 1; q < S_; q++) yi = (sg - (Tab::C(q) - 1.0)) * (yi + S.cont[(size_t)((S_ - 1 - q) * MN + e) * M + i]);
      yout[s.nout * MN + e] = yi + S.y[(size_t)e * M + i];
    }
    tout[s.nout] = tn;
    s.nout += 1;
  }
}

// radau.m:1267-1269: the last recorded point is overwritten with the final (t, y)
JB_HD void dense_finish_g(const State& S, const TrajS& s, long long i) {
  if (s.nout >= 1) {
    S.tout[(size_t)i * S.ntspan + s.nout - 1] = s.t;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
    for (int e = 0; e < MN; e++) S.yout[((size_t)i * S.ntspan + s.nout - 1) * MN + e] = S.y[(size_t)e * S.M + i];
  }
}

// new step size + accept / reject up to the dynamics evaluation at the new point (radau.m:1038-1253)
template <int S_>
JB_HD void controller_a(const State& S, const RadauParams& P, TrajS& s, long long i) {
  using Tab = RadauTab<S_>;
  const long long M = S.M;
  const int Nit = nit_of(S_, P.MaxNbrNewton);
  const double err = s.err;
  if (bad(err)) {
    s.res = 2;
    finish(s);
    return;
  }
  const double fac = fmin(P.Safe, (2.0 * Nit + 1.0) / (2.0 * Nit + s.Newt));
  double quot = fmax(P.FacR, fmin(P.FacL, det_pow(err, 1.0 / S_ + 1.0) / fac));
  double hnew = s.h / quot;
  if (err < 1.0) {
    s.flags &= ~PF_FIRST;
    s.st[ST_ACC] += 1;
    const bool gus_upd = (P.Gustafsson & 1) && !JB_FLAG(s, PF_CHANGE);   // see radau_core.cuh (quirk Q8)
    const bool gus_lim = (P.Gustafsson & 2) ? true : gus_upd;
    if (gus_lim) {
      if (s.st[ST_ACC] > 1) {
        const double ratio = err * err / s.erracc;
        double root;
        if (S_ == 1) root = ::sqrt(ratio);
        else if (S_ == 3) root = ::sqrt(::sqrt(ratio));
        else if (S_ == 7) root = ::sqrt(::sqrt(::sqrt(ratio)));
        else root = det_pow(ratio, 1.0 / (S_ + 1.0));
        double facgus = (s.hacc / s.h) * root / P.Safe;
        facgus = fmax(P.FacR, fmin(P.FacL, facgus));
        quot = fmax(quot, facgus);
        hnew = s.h / quot;
      }
    }
    if (gus_upd) {
      s.hacc = s.h;
      s.erracc = fmax(1.0e-2, err);
    }
    s.h_old = s.h;
    s.t = s.t + s.h;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
    for (int e = 0; e < MN; e++) {
      double zz[S_], cc[S_];
#pragma unroll
      for (int q = 0; q < S_; q++) zz[q] = S.z[(size_t)(q * MN + e) * M + i];
      S.y[(size_t)e * M + i] = S.y[(size_t)e * M + i] + zz[S_ - 1];
      cc[S_ - 1] = zz[0] * (1.0 / Tab::C(0));
#pragma unroll
      for (int q = 1; q <= S_ - 1; q++) {
        const double Fact = 1.0 / (Tab::C(S_ - q - 1) - Tab::C(S_ - q));
        cc[q - 1] = (zz[S_ - q - 1] - zz[S_ - q]) * Fact;
      }
#pragma unroll
      for (int jj = 2; jj <= S_; jj++)
#pragma unroll
        for (int k = S_; k >= jj; k--) {
          const double Fact = (S_ - k == 0) ? 1.0 / (-Tab::C(jj - 1)) : 1.0 / (Tab::C(S_ - k - 1) - Tab::C(S_ - k + jj - 1));
          cc[k - 1] = (cc[k - 1] - cc[k - 2]) * Fact;
        }
#pragma unroll
      for (int q = 0; q < S_; q++) S.cont[(size_t)(q * MN + e) * M + i] = cc[q];
    }
    if (S.tout) dense_output_g<S_>(S, s, i);   // radau.m:1150-1168
    s.flags |= PF_NEED_JAC;
    if (JB_FLAG(s, PF_LAST)) {
      s.h = s.hopt;
      s.st[ST_REJ] += 1;
      finish(s);
      return;
    }
    s.st[ST_FCN] += 1;
    s.hnew = hnew;
    s.phase = PH_F0_EVAL;
  } else {
    s.flags |= PF_REJECT;
    s.flags &= ~PF_LAST;
    if (JB_FLAG(s, PF_FIRST)) s.h = s.h / 10.0;
    else s.h = hnew;
    if (s.st[ST_ACC] >= 1) s.st[ST_REJ] += 1;
    s.flags |= PF_NEED_LU;
    s.phase = PH_IDLE;
  }
}

JB_HD double scaled_err(const State& S, const RadauParams& P, long long i) {
  const long long M = S.M;
  double sum = 0.0;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) {
    const double v = S.ev[(size_t)e * M + i] * S.rscal[(size_t)e * M + i];
    sum = fma(v, v, sum);
  }
  return fmax(::sqrt(sum) * P.inv_sqrt_n, 1e-10);
}

// Estrad, first solve (radaute_impl.hpp:1902-1948)
template <int S_>
JB_HD void estrad_a(const State& S, const RadauParams& P, TrajS& s, long long i) {
  using Tab = RadauTab<S_>;
  const long long M = S.M;
  const double rh = 1.0 / s.h;
  s.work += W_tail(S_, MN, (int)FLOPS_F);
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) {
    double sum = S.z[(size_t)e * M + i] * (Tab::Dd(0) * rh);
#pragma unroll
    for (int q = 1; q < S_; q++) sum = fma(S.z[(size_t)(q * MN + e) * M + i], Tab::Dd(q) * rh, sum);
    S.temp[(size_t)e * M + i] = sum;
    S.ev[(size_t)e * M + i] = S.f0[(size_t)e * M + i] + sum;
  }
  rlu_solve_g(S.L0a(i), S.L0rd(i), S.L0piv(i), 1, M, S.ev + i);
  s.err = scaled_err(S, P, i);
  if (s.err >= 1.0 && (JB_FLAG(s, PF_FIRST) || JB_FLAG(s, PF_REJECT))) {
    S.errs[i] = s.err;   // sic: the scalar err is added to every component (radaute_impl.hpp:1951)
    s.st[ST_FCN] += 1;
    s.work += FLOPS_F + 2LL * MN * MN + 3LL * MN;
    s.phase = PH_ERR_EVAL;
    return;
  }
  controller_a<S_>(S, P, s, i);
}

template <int S_>
JB_HD void estrad_b(const State& S, const RadauParams& P, TrajS& s, long long i) {
  const long long M = S.M;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) S.ev[(size_t)e * M + i] = S.ev[(size_t)e * M + i] + S.temp[(size_t)e * M + i];
  rlu_solve_g(S.L0a(i), S.L0rd(i), S.L0piv(i), 1, M, S.ev + i);
  s.err = scaled_err(S, P, i);
  controller_a<S_>(S, P, s, i);
}

JB_HD void ph_estrad_b(const State& S, const RadauParams& P, long long i) {
  TrajS s = S.sc[i];
  if (s.phase != PH_ERR_EVAL) return;
  switch (s.NbrStg) {
    case 1: estrad_b<1>(S, P, s, i); break;
    case 3: estrad_b<3>(S, P, s, i); break;
    case 5: estrad_b<5>(S, P, s, i); break;
    default: estrad_b<7>(S, P, s, i); break;
  }
  S.sc[i] = s;
}

// one simplified-Newton iteration after its stage evaluations (radau.m:932-1018), in three parts so
// that the GPU can run the linear solves with 16 lanes per trajectory (lu16.cuh) between them:
//   newton_pre    z = TI f, right-hand sides of Solvrad        (false: NaN, trajectory finished)
//   newton_solves the real + complex triangular solves          (scalar form; GPU: solve16_kernel)
//   newton_post   norms, convergence logic, w += z, z = T w, then Estrad / controller or next iteration
template <int S_>
JB_HD bool newton_pre(const State& S, const RadauParams& P, TrajS& s, long long i) {
  using Tab = RadauTab<S_>;
  const long long M = S.M;
  const double rh = 1.0 / s.h;
  bool nanf = false;
  for (int k = 0; k < S_ * MN; k++) {
    const double v = S.F[(size_t)k * M + i];
    nanf = nanf || (v != v);
  }
  if (nanf) {
    s.res = 2;
    finish(s);
    return false;
  }
  s.st[ST_FCN] += S_;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) {
    double fq[S_];
#pragma unroll
    for (int q = 0; q < S_; q++) fq[q] = S.F[(size_t)(q * MN + e) * M + i];
#pragma unroll
    for (int j = 0; j < S_; j++) {
      double acc = Tab::TI(j, 0) * fq[0];
#pragma unroll
      for (int q = 1; q < S_; q++) acc = fma(Tab::TI(j, q), fq[q], acc);
      S.z[(size_t)(j * MN + e) * M + i] = acc;
    }
  }
  {  // Solvrad right-hand sides
    const double valp0 = Tab::ValP(0) * rh;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
    for (int e = 0; e < MN; e++) S.z[(size_t)e * M + i] = fma(-valp0, S.w[(size_t)e * M + i], S.z[(size_t)e * M + i]);
#pragma unroll
    for (int q = 1; q <= (S_ - 1) / 2; q++) {
      const int q2 = 2 * q - 1, q3 = 2 * q;
      const double v2 = Tab::ValP(q2) * rh, v3 = Tab::ValP(q3) * rh;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
      for (int e = 0; e < MN; e++) {
        const double w2 = S.w[(size_t)(q2 * MN + e) * M + i], w3 = S.w[(size_t)(q3 * MN + e) * M + i];
        const double a = fma(v3, w3, fma(-v2, w2, S.z[(size_t)(q2 * MN + e) * M + i]));
        const double b = fma(-v2, w3, fma(-v3, w2, S.z[(size_t)(q3 * MN + e) * M + i]));
        S.z[(size_t)(q2 * MN + e) * M + i] = a;
        S.z[(size_t)(q3 * MN + e) * M + i] = b;
      }
    }
  }
  return true;
}

template <int S_>
JB_HD void newton_solves(const State& S, long long i) {
  const long long M = S.M;
  rlu_solve_g(S.L0a(i), S.L0rd(i), S.L0piv(i), 1, M, S.z + i);
#pragma unroll
  for (int q = 1; q <= (S_ - 1) / 2; q++)
    zlu_solve_g(S.Lre(i, q - 1), S.Lim(i, q - 1), S.Lrdre(i, q - 1), S.Lrdim(i, q - 1), S.Lpiv(i, q - 1), 1, M,
                S.z + (size_t)(2 * q - 1) * MN * M + i, S.z + (size_t)(2 * q) * MN * M + i);
}

template <int S_>
JB_HD void newton_post(const State& S, const RadauParams& P, TrajS& s, long long i) {
  using Tab = RadauTab<S_>;
  constexpr int si = (S_ - 1) / 2;
  const long long M = S.M;
  const int Nit = nit_of(S_, P.MaxNbrNewton);
  const double FNewt = P.FNewt[si];
  s.st[ST_SOL] += 1;
  s.work += W_newt(S_, MN, (int)FLOPS_F);
  double NewNrm = 0.0;
#pragma unroll
  for (int q = 0; q < S_; q++) {
    double sum = 0.0;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
    for (int e = 0; e < MN; e++) {
      const double v = S.z[(size_t)(q * MN + e) * M + i] * S.rscal[(size_t)e * M + i];
      sum = fma(v, v, sum);
    }
    NewNrm += ::sqrt(sum);
  }
  NewNrm = NewNrm * P.inv_sqrt_sn[si];
  s.NewNrm = NewNrm;
  if (bad(NewNrm)) {
    s.res = 2;
    finish(s);
    return;
  }
  if (s.Newt > 1 && s.Newt < Nit) {
    const double thq = NewNrm / s.OldNrm;
    if (s.Newt == 2) s.Theta = thq;
    else s.Theta = ::sqrt(thq * s.thqold);
    s.thqold = thq;
    if (s.Theta < 0.99) {
      s.FacConv = s.Theta / (1.0 - s.Theta);
      const double dyth = s.FacConv * NewNrm * ipow(s.Theta, Nit - 1 - s.Newt) / FNewt;
      if (dyth >= 1.0) {
        const double qNewt = fmax(1.0e-4, fmin(20.0, dyth));
        const double hhfac = 0.8 * det_pow(qNewt, -1.0 / (4.0 + Nit - 1 - s.Newt));
        s.h = hhfac * s.h;
        s.flags |= PF_REJECT | PF_NEED_LU;
        s.flags &= ~PF_LAST;
        set_flag(s, PF_UNEXP_NEWT, hhfac <= 0.5);
        s.st[ST_NREJ] += 1;
        s.st[ST_REJ] += 1;
        s.phase = PH_IDLE;
        return;
      }
    } else {
      s.h = s.h * 0.5;
      s.flags |= PF_REJECT | PF_UNEXP_STEP | PF_NEED_LU;
      s.flags &= ~PF_LAST;
      s.st[ST_REJ] += 1;
      s.phase = PH_IDLE;
      return;
    }
  }
  s.OldNrm = fmax(NewNrm, EPS);
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) {
    double ww[S_];
#pragma unroll
    for (int q = 0; q < S_; q++) {
      ww[q] = S.w[(size_t)(q * MN + e) * M + i] + S.z[(size_t)(q * MN + e) * M + i];
      S.w[(size_t)(q * MN + e) * M + i] = ww[q];
    }
#pragma unroll
    for (int j = 0; j < S_; j++) {
      double acc = Tab::T(j, 0) * ww[0];
#pragma unroll
      for (int q = 1; q < S_; q++) acc = fma(Tab::T(j, q), ww[q], acc);
      S.z[(size_t)(j * MN + e) * M + i] = acc;
    }
  }
  if (s.FacConv * NewNrm > FNewt) {
    newton_begin<S_>(s, P);
    return;
  }
  estrad_a<S_>(S, P, s, i);
}

template <int S_>
JB_HD void newton_step(const State& S, const RadauParams& P, TrajS& s, long long i) {
  if (!newton_pre<S_>(S, P, s, i)) return;
  newton_solves<S_>(S, i);
  newton_post<S_>(S, P, s, i);
}

JB_HD void ph_newton(const State& S, const RadauParams& P, long long i) {
  TrajS s = S.sc[i];
  if (s.phase != PH_NEWTON) return;
  switch (s.NbrStg) {
    case 1: newton_step<1>(S, P, s, i); break;
    case 3: newton_step<3>(S, P, s, i); break;
    case 5: newton_step<5>(S, P, s, i); break;
    default: newton_step<7>(S, P, s, i); break;
  }
  S.sc[i] = s;
}

// the GPU's split of ph_newton around solve16_kernel
JB_HD void ph_newton_pre(const State& S, const RadauParams& P, long long i) {
  TrajS s = S.sc[i];
  if (s.phase != PH_NEWTON) return;
  bool ok;
  switch (s.NbrStg) {
    case 1: ok = newton_pre<1>(S, P, s, i); break;
    case 3: ok = newton_pre<3>(S, P, s, i); break;
    case 5: ok = newton_pre<5>(S, P, s, i); break;
    default: ok = newton_pre<7>(S, P, s, i); break;
  }
  (void)ok;
  S.sc[i] = s;
}
JB_HD void ph_newton_post(const State& S, const RadauParams& P, long long i) {
  TrajS s = S.sc[i];
  if (s.phase != PH_NEWTON) return;   // finished by ph_newton_pre (NaN in the stage values)
  switch (s.NbrStg) {
    case 1: newton_post<1>(S, P, s, i); break;
    case 3: newton_post<3>(S, P, s, i); break;
    case 5: newton_post<5>(S, P, s, i); break;
    default: newton_post<7>(S, P, s, i); break;
  }
  S.sc[i] = s;
}

// after the dynamics evaluation at the accepted point (radau.m:1199-1232)
JB_HD void ph_accept_b(const State& S, const RadauParams& P, long long i) {
  const long long M = S.M;
  TrajS s = S.sc[i];
  if (s.phase != PH_F0_EVAL) return;
  bool nanf = false;
#pragma unroll 8   // 8 components in flight: the loads of a one-thread-per-trajectory phase kernel must overlap
  for (int e = 0; e < MN; e++) {
    const double v = S.f0[(size_t)e * M + i];
    nanf = nanf || (v != v);
  }
  if (nanf) {
    s.res = 2;
    finish(s);
    S.sc[i] = s;
    return;
  }
  double hnew = s.hnew;
  hnew = s.PosNeg * fmin(fabs(hnew), fabs(s.hmaxn));
  s.hopt = s.PosNeg * fmin(fabs(s.h), fabs(hnew));
  if (JB_FLAG(s, PF_REJECT)) hnew = s.PosNeg * fmin(fabs(hnew), fabs(s.h));
  s.flags &= ~PF_REJECT;
  s.phase = PH_IDLE;
  if ((s.t + hnew / P.Quot1 - s.tfinal) * s.PosNeg >= 0) {
    s.h = s.tfinal - s.t;
    s.flags |= PF_LAST;
  } else {
    const double qt = hnew / s.h;
    if (s.Theta <= P.Thet && qt >= P.Quot1 && qt <= P.Quot2) {
      s.flags |= PF_KEEP;
      s.flags &= ~(PF_NEED_JAC | PF_NEED_LU);
      S.sc[i] = s;
      return;
    }
    s.h = hnew;
  }
  s.flags |= PF_NEED_LU;
  if (s.Theta <= P.Thet) s.flags &= ~PF_NEED_JAC;
  S.sc[i] = s;
}

}  // namespace mlp
}  // namespace jb
