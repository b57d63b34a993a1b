// TEST TOOL ONLY: runs the integrator core (janus_ode_b200/csrc/radau_core.cuh) on the CPU.
// The product library (libjanus_b200.so) never contains this; it exists so that the
// arithmetic contract (GPU core == oracle "mirror" mode, bit for bit) can be checked in the
// GPU-less container.  Built by tests/hostsim/build.sh with g++ -ffp-contract=off.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../janus_ode_b200/csrc/radau_core.cuh"
#include "../../janus_ode_b200/csrc/seulex.cuh"
#include "../../janus_ode_b200/csrc/dopri853.cuh"

using namespace jb;

// dense output request for the whole batch (tout == nullptr: none); tspan is [M][g_nts]
static int g_nts = 2;
static double *g_tout = nullptr, *g_yout = nullptr;
static int* g_nout = nullptr;

template <class Ham, int SM, bool VAR>
static void run(const RadauParams& P, const double* y0, const double* tspan, long long M, double* y_out,
                double* t_out, double* h_out, int* res, long long* stats, int* nstg) {
  constexpr int NY = Ham::NV;
  for (long long m = 0; m < M; m++) {
    Traj<NY, SM> T;
    traj_init<Ham, NY, SM>(T, P, y0 + m * NY, tspan[g_nts * m], tspan[g_nts * m + g_nts - 1]);
    DenseOut d{tspan + m * g_nts, g_nts, g_tout ? g_tout + m * g_nts : nullptr, g_yout ? g_yout + m * g_nts * NY : nullptr};
    integrate<Ham, NY, SM, VAR>(T, P, g_tout ? &d : nullptr);
    if (g_tout) {
      dense_finish<NY, SM>(T, d);
      g_nout[m] = T.nout;
    }
    for (int i = 0; i < NY; i++) y_out[m * NY + i] = T.y[i];
    t_out[m] = T.t;
    h_out[m] = T.h;
    res[m] = T.res;
    for (int k = 0; k < 8; k++) stats[m * 8 + k] = T.st[k];
    nstg[m] = T.NbrStg;
  }
}

template <class Ham>
static void dispatch(const RadauParams& P, const double* y0, const double* tspan, long long M, double* y_out,
                     double* t_out, double* h_out, int* res, long long* stats, int* nstg) {
  if (P.MinNbrStg == P.MaxNbrStg && P.MaxNbrStg == 3) run<Ham, 3, false>(P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg);
  else if (P.MinNbrStg == P.MaxNbrStg && P.MaxNbrStg == 1) run<Ham, 1, false>(P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg);
  else if (P.MaxNbrStg <= 3) run<Ham, 3, true>(P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg);
  else run<Ham, 7, true>(P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg);
}

template <class Ham>
static void seulex_run(const double* par, int sign, long long M, double atol, double rtol, int KMAXX, int quirks,
                       const double* htry, double* y, double* x, double* theta, int* k_targ, double* hnext, double* hdid,
                       int* flags, double* yjac, int* k_out, int* nrej) {
  constexpr int NY = Ham::NV;
  for (long long m = 0; m < M; m++)
    sx::seulex_step_one<Ham>(par, (double)sign, atol, rtol, KMAXX, quirks, htry[m], y + m * NY, x[m], theta[m], k_targ[m],
                             hnext[m], hdid[m], flags[m], yjac + m * NY, k_out[m], nrej[m]);
}

extern "C" {
// the SEULEX step of csrc/seulex.cuh on the CPU (same source as the kernel)
int hostsim_seulex_step(int ham_id, const double* par, int sign, long long M, double atol, double rtol, int KMAXX, int quirks,
                        const double* htry, double* y, double* x, double* theta, int* k_targ, double* hnext, double* hdid,
                        int* flags, double* yjac, int* k_out, int* nrej) {
  switch (ham_id) {
    case HAM_HARMONIC: seulex_run<HamHarmonic>(par, sign, M, atol, rtol, KMAXX, quirks, htry, y, x, theta, k_targ, hnext, hdid, flags, yjac, k_out, nrej); return 0;
    case HAM_VDP: seulex_run<HamVdp>(par, sign, M, atol, rtol, KMAXX, quirks, htry, y, x, theta, k_targ, hnext, hdid, flags, yjac, k_out, nrej); return 0;
    case HAM_VDP_PMP_SMOOTH: seulex_run<HamVdpPmp<false>>(par, sign, M, atol, rtol, KMAXX, quirks, htry, y, x, theta, k_targ, hnext, hdid, flags, yjac, k_out, nrej); return 0;
    case HAM_VDP_PMP_BANG: seulex_run<HamVdpPmp<true>>(par, sign, M, atol, rtol, KMAXX, quirks, htry, y, x, theta, k_targ, hnext, hdid, flags, yjac, k_out, nrej); return 0;
  }
  return -1;
}

void hostsim_set_dense(int nts, double* tout, double* yout, int* nout) {
  g_nts = nts; g_tout = tout; g_yout = yout; g_nout = nout;
}
// P is passed as raw bytes of jb::RadauParams (filled by the caller, see tests/hostsim.py)
int hostsim_params_size() { return (int)sizeof(RadauParams); }

void hostsim_fill_params(RadauParams* P, const double* RelTol, const double* AbsTol, int ny, double InitialStep,
                         double Thet, double Safe, double Quot1, double Quot2, double FacL, double FacR, double Vitu,
                         double Vitd, double hhou, double hhod, int sign, long long MaxNbrStep, int Start_Newt,
                         int MaxNbrNewton, int NbrStg, int MinNbrStg, int MaxNbrStg, int Gustafsson, const double* hp,
                         int nhp) {
  std::memset(P, 0, sizeof *P);
  for (int k = 0; k < 4; k++) {
    const int s = 2 * k + 1;
    const double ExpmNs = (s + 1.0) / (2.0 * s);
    double mrt = INFINITY;
    for (int i = 0; i < ny; i++) {
      const double quot = AbsTol[i] / RelTol[i];
      P->RelTol1[k][i] = 0.1 * std::pow(RelTol[i], ExpmNs);
      P->AbsTol1[k][i] = P->RelTol1[k][i] * quot;
      mrt = std::fmin(mrt, P->RelTol1[k][i]);
    }
    P->FNewt[k] = std::fmax(10 * EPS / mrt, std::fmin(0.03, std::pow(mrt, 1.0 / ExpmNs - 1.0)));
    if (s == 1) P->FNewt[k] = std::fmax(10 * EPS / mrt, 0.03);
    P->inv_sqrt_sn[k] = 1.0 / std::sqrt((double)(s * ny));
  }
  P->inv_sqrt_n = 1.0 / std::sqrt((double)ny);
  P->InitialStep = InitialStep; P->Thet = Thet; P->Safe = Safe; P->Quot1 = Quot1; P->Quot2 = Quot2;
  P->FacL = 1.0 / FacL; P->FacR = 1.0 / FacR; P->Vitu = Vitu; P->Vitd = Vitd; P->hhou = hhou; P->hhod = hhod;
  P->sign = (double)sign; P->MaxNbrStep = MaxNbrStep; P->Start_Newt = Start_Newt; P->MaxNbrNewton = MaxNbrNewton;
  P->NbrStg = NbrStg; P->MinNbrStg = MinNbrStg; P->MaxNbrStg = MaxNbrStg; P->Gustafsson = Gustafsson;
  for (int i = 0; i < nhp && i < 8; i++) P->ham_par[i] = hp[i];
}

int hostsim_solve(int ham_id, const RadauParams* P, const double* y0, const double* tspan, long long M, double* y_out,
                  double* t_out, double* h_out, int* res, long long* stats, int* nstg) {
  switch (ham_id) {
    case HAM_HARMONIC: dispatch<HamHarmonic>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg); return 0;
    case HAM_VDP: dispatch<HamVdp>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg); return 0;
    case HAM_VDP_PMP_SMOOTH: dispatch<HamVdpPmp<false>>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg); return 0;
    case HAM_VDP_PMP_BANG: dispatch<HamVdpPmp<true>>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg); return 0;
    case HAM_VDP_AUG_BANG: dispatch<HamVdpAugBang>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg); return 0;
  }
  return -1;
}
}

// ---- epoch emulation: init -> park; repeat { bucket by order; run_epoch<S> for <= K passes; park }
template <class Ham>
static void run_epochs(const RadauParams& P, const double* y0, const double* tspan, long long M, double* y_out,
                       double* t_out, double* h_out, int* res, long long* stats, int* nstg, int K) {
  const bool stop = K < 0;   // K < 0: incremental step() protocol, one accepted step per sweep
  if (stop) K = 1 << 30;
  bool stepped = false;
  constexpr int NY = Ham::NV;
  std::vector<Parked<NY>> st(M);
  for (long long m = 0; m < M; m++) {
    Traj<NY, 7> T;
    traj_init<Ham, NY, 7>(T, P, y0 + m * NY, tspan[2 * m], tspan[2 * m + 1]);
    park<NY, 7>(T, tspan[2 * m], false, false, st[m]);
  }
  for (int epoch = 0; epoch < 100000; epoch++) {
    long long active = 0;
    for (long long m = 0; m < M; m++) {
      Parked<NY>& p = st[m];
      if (p.flags & PF_FINISHED) continue;
      active++;
      double t0;
      bool hd, fin;
      switch (p.NbrStg) {
        case 1: { Traj<NY, 1> T; unpark<NY, 1>(p, T, t0, hd); fin = run_epoch<Ham, NY, 1>(T, P, hd, K, nullptr, stop, &stepped); park<NY, 1>(T, t0, hd, fin, p); } break;
        case 3: { Traj<NY, 3> T; unpark<NY, 3>(p, T, t0, hd); fin = run_epoch<Ham, NY, 3>(T, P, hd, K, nullptr, stop, &stepped); park<NY, 3>(T, t0, hd, fin, p); } break;
        case 5: { Traj<NY, 5> T; unpark<NY, 5>(p, T, t0, hd); fin = run_epoch<Ham, NY, 5>(T, P, hd, K, nullptr, stop, &stepped); park<NY, 5>(T, t0, hd, fin, p); } break;
        default: { Traj<NY, 7> T; unpark<NY, 7>(p, T, t0, hd); fin = run_epoch<Ham, NY, 7>(T, P, hd, K, nullptr, stop, &stepped); park<NY, 7>(T, t0, hd, fin, p); } break;
      }
    }
    if (!active) break;
  }
  for (long long m = 0; m < M; m++) {
    const Parked<NY>& p = st[m];
    for (int i = 0; i < NY; i++) y_out[m * NY + i] = p.y[i];
    t_out[m] = p.t; h_out[m] = p.h; res[m] = p.res; nstg[m] = p.NbrStg;
    for (int k = 0; k < 8; k++) stats[m * 8 + k] = p.st[k];
  }
}

extern "C" int hostsim_solve_epochs(int ham_id, const RadauParams* P, const double* y0, const double* tspan, long long M,
                                    double* y_out, double* t_out, double* h_out, int* res, long long* stats, int* nstg,
                                    int K) {
  switch (ham_id) {
    case HAM_HARMONIC: run_epochs<HamHarmonic>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, K); return 0;
    case HAM_VDP: run_epochs<HamVdp>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, K); return 0;
    case HAM_VDP_PMP_SMOOTH: run_epochs<HamVdpPmp<false>>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, K); return 0;
    case HAM_VDP_PMP_BANG: run_epochs<HamVdpPmp<true>>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, K); return 0;
    case HAM_VDP_AUG_BANG: run_epochs<HamVdpAugBang>(*P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, K); return 0;
  }
  return -1;
}


// ---- forward sensitivities (radau_core.cuh Sens / sens_step) on the CPU: the all-orders instantiation the GPU's
// sens_kernel<Ham, 7, true> runs
template <class Ham>
static void run_sens(const RadauParams& P, const double* y0, const double* tspan, long long M, double* y_out, double* S_out,
                     long long* stats, int* iters, double* Sp_out = nullptr) {
  constexpr int NY = Ham::NV;
  for (long long m = 0; m < M; m++) {
    Traj<NY, 7> T;
    traj_init<Ham, NY, 7>(T, P, y0 + m * NY, tspan[2 * m], tspan[2 * m + 1]);
    Sens<NY> sn;
    sn.iters = 0;
    sn.unconverged = 0;
    for (int k = 0; k < NY * NY; k++) sn.S[k] = (k / NY == k % NY) ? 1.0 : 0.0;
    for (int k = 0; k < P.sens_npar * NY; k++) sn.Sp[k] = 0.0;
    integrate<Ham, NY, 7, true>(T, P, nullptr, &sn);
    for (int k = 0; k < P.sens_npar * NY; k++) Sp_out[m * P.sens_npar * NY + k] = sn.Sp[k];
    for (int i = 0; i < NY; i++) y_out[m * NY + i] = T.y[i];
    for (int k = 0; k < NY * NY; k++) S_out[m * NY * NY + k] = sn.S[k];
    for (int k = 0; k < 8; k++) stats[m * 8 + k] = T.st[k];
    iters[m] = sn.iters;
  }
}

extern "C" int hostsim_solve_sens(int ham_id, const RadauParams* P, const double* y0, const double* tspan, long long M,
                                  double* y_out, double* S_out, long long* stats, int* iters) {
  switch (ham_id) {
    case HAM_HARMONIC: run_sens<HamHarmonic>(*P, y0, tspan, M, y_out, S_out, stats, iters); return 0;
    case HAM_VDP: run_sens<HamVdp>(*P, y0, tspan, M, y_out, S_out, stats, iters); return 0;
  }
  return -1;
}

// ---- explicit DOPRI853 with a dual part (csrc/dopri853.cuh dopri_step_one + dopri_dual_one), integrated to xf on the host
template <class Ham>
static void run_dopri_dual(const double* par, double sign, const double* y0, long long M, double x0, double xf, double atol,
                           double rtol, double h0, double* y_out, double* S_out, int* steps) {
  constexpr int NY = Ham::NV;
  for (long long m = 0; m < M; m++) {
    double y[NY], d[NY], Sd[NY * NY], x = x0, errold = 1.0e-4, hnext = h0, hdid = 0.0;
    for (int i = 0; i < NY; i++) y[i] = y0[m * NY + i];
    for (int k = 0; k < NY * NY; k++) Sd[k] = (k / NY == k % NY) ? 1.0 : 0.0;
    ham_dyns<Ham>(y, par, sign, d);
    int n = 0, nrej;
    bool aliased = false;
    while (xf - x > 1e-14 * (1.0 + fabs(xf)) && n < 100000) {
      const double htry = h0 < xf - x ? h0 : xf - x;   // the caller's htry every call (the reference's hnext never grows)
      double yb[NY], k1[NY];
      for (int i = 0; i < NY; i++) k1[i] = d[i];
      const double xb = x;
      dp::dopri_step_one<Ham>(par, sign, M, atol, rtol, aliased, x, y, d, errold, hnext, htry, hdid, nrej, yb);
      if (!(hdid == hdid) || x == xb) break;
      if (S_out) dp::dopri_dual_one<Ham>(par, sign, yb, k1, hdid, Sd, NY);
      n++;   // aliased stays false: a clean stepper (the reference's y / yout aliasing corrupts the state after a rejection)
    }
    for (int i = 0; i < NY; i++) y_out[m * NY + i] = y[i];
    if (S_out) for (int k = 0; k < NY * NY; k++) S_out[m * NY * NY + k] = Sd[k];
    steps[m] = n;
  }
}
extern "C" int hostsim_dopri_dual(int ham_id, const double* par, double sign, const double* y0, long long M, double x0, double xf,
                                  double atol, double rtol, double h0, double* y_out, double* S_out, int* steps) {
  switch (ham_id) {
    case HAM_HARMONIC: run_dopri_dual<HamHarmonic>(par, sign, y0, M, x0, xf, atol, rtol, h0, y_out, S_out, steps); return 0;
    case HAM_VDP: run_dopri_dual<HamVdp>(par, sign, y0, M, x0, xf, atol, rtol, h0, y_out, S_out, steps); return 0;
  }
  return -1;
}

// ---- SEULEX with a dual part (csrc/seulex.cuh seulex_step_one + seulex_dual_one), integrated to xf on the host
template <class Ham>
static void run_seulex_dual(const double* par, double sign, const double* y0, long long M, double x0, double xf, double atol,
                            double rtol, int KMAXX, double h0, double* y_out, double* S_out, int* steps) {
  constexpr int NY = Ham::NV;
  const double jac_redo = std::fmin(1.0e-4, rtol);
  const double logfact = -std::log10(rtol + atol) * 0.6 + 0.5;
  const int kt0 = (int)std::floor(std::fmax(1.0, std::fmin((double)(KMAXX - 1), std::nearbyint(logfact))));
  for (long long m = 0; m < M; m++) {
    double y[NY], yj[NY], Sd[NY * NY], Sj[NY * NY], x = x0, theta = 2.0 * jac_redo, hnext = -1.0e99, hdid = 0.0;
    int k_targ = kt0, flags = sx::F_FIRST, k = 0, nrej = 0, n = 0;
    for (int i = 0; i < NY; i++) y[i] = yj[i] = y0[m * NY + i];
    for (int e = 0; e < NY * NY; e++) Sd[e] = Sj[e] = (e / NY == e % NY) ? 1.0 : 0.0;
    double htry = h0;
    while (xf - x > 1e-14 * (1.0 + std::fabs(xf)) && n < 100000) {
      if (htry > xf - x) htry = xf - x;
      double ysav[NY], h_taken = 0.0;
      bool fresh = false;
      for (int i = 0; i < NY; i++) ysav[i] = y[i];
      sx::seulex_step_one<Ham>(par, sign, atol, rtol, KMAXX, 0, htry, y, x, theta, k_targ, hnext, hdid, flags, yj, k, nrej, &fresh,
                               &h_taken);
      if (S_out) {
        if (fresh) for (int e = 0; e < NY * NY; e++) Sj[e] = Sd[e];
        if (nrej <= 200 && h_taken != 0.0) sx::seulex_dual_one<Ham>(par, sign, ysav, yj, Sj, h_taken, k, Sd, NY);
      }
      htry = hnext;
      n++;
    }
    for (int i = 0; i < NY; i++) y_out[m * NY + i] = y[i];
    if (S_out) for (int e = 0; e < NY * NY; e++) S_out[m * NY * NY + e] = Sd[e];
    steps[m] = n;
  }
}
extern "C" int hostsim_seulex_dual(int ham_id, const double* par, double sign, const double* y0, long long M, double x0, double xf,
                                   double atol, double rtol, int KMAXX, double h0, double* y_out, double* S_out, int* steps) {
  switch (ham_id) {
    case HAM_HARMONIC: run_seulex_dual<HamHarmonic>(par, sign, y0, M, x0, xf, atol, rtol, KMAXX, h0, y_out, S_out, steps); return 0;
    case HAM_VDP: run_seulex_dual<HamVdp>(par, sign, y0, M, x0, xf, atol, rtol, KMAXX, h0, y_out, S_out, steps); return 0;
  }
  return -1;
}

// the device's deterministic tanh (csrc/detmath.cuh det_tanh), run on the host
extern "C" void hostsim_det_tanh(const double* x, long long n, double* out) {
  for (long long i = 0; i < n; i++) out[i] = det_tanh(x[i]);
}

// with parameter directions d y / d ham_par[par_index[k]] (Sp_out [M, n_par, NY])
extern "C" int hostsim_solve_sens_par(int ham_id, const RadauParams* P0, const double* y0, const double* tspan, long long M,
                                      int n_par, const int* par_index, double* y_out, double* S_out, double* Sp_out,
                                      long long* stats, int* iters) {
  RadauParams P = *P0;
  P.sens_npar = n_par;
  for (int k = 0; k < n_par; k++) P.sens_par[k] = par_index[k];
  switch (ham_id) {
    case HAM_HARMONIC: run_sens<HamHarmonic>(P, y0, tspan, M, y_out, S_out, stats, iters, Sp_out); return 0;
    case HAM_VDP: run_sens<HamVdp>(P, y0, tspan, M, y_out, S_out, stats, iters, Sp_out); return 0;
    case HAM_VDP_PMP_SMOOTH: run_sens<HamVdpPmp<false>>(P, y0, tspan, M, y_out, S_out, stats, iters, Sp_out); return 0;
  }
  return -1;
}

// ---- MLP Hamiltonian path (csrc/mlp_phase.cuh): the same phase functions the GPU kernels wrap,
// driven by the same pass protocol; the stage / Jacobian evaluations (tile GEMMs on the GPU) are
// delegated to callbacks (the mirror oracle's scalar evaluation, passed in from the test).
#include "../../janus_ode_b200/csrc/mlp_phase.cuh"

typedef int (*mlp_dyns_fn)(int problem, const double* par, int sign, int Ny, long long M, const double* y, double* f);
typedef int (*mlp_jac_fn)(int problem, const double* par, int sign, int Ny, long long M, const double* y, double* J);

extern "C" int hostsim_mlp_solve(mlp_dyns_fn dyns, mlp_jac_fn jacf, const double* par, const RadauParams* Pp,
                                 const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                                 double* h_out, int* res, long long* stats, int* nstg, long long* work) {
  using namespace jb::mlp;
  const RadauParams& P = *Pp;
  const int sign = (int)P.sign, SM = P.MaxNbrStg, nq = (SM - 1) / 2 > 0 ? (SM - 1) / 2 : 1;
  std::vector<TrajS> sc(M);
  auto dv = [&](size_t per) { return std::vector<double>(per * M, 0.0); };
  (void)nq;
  auto y = dv(MN), f0 = dv(MN), Jac = dv(MN * MN), cont = dv(SM * MN), z = dv(SM * MN), w = dv(SM * MN), F = dv(SM * MN),
       rscal = dv(MN), temp = dv(MN), ev = dv(MN), errs = dv(1), LU = dv(lu_doubles(SM));
  std::vector<int> LUpiv((size_t)lu_ints(SM) * M);
  State S{M, SM, sc.data(), y.data(), f0.data(), cont.data(), z.data(), w.data(), F.data(), rscal.data(),
          temp.data(), ev.data(), errs.data(), Jac.data(), LU.data(), LUpiv.data(), lu_doubles(SM), lu_ints(SM),
          2, tspan, nullptr, nullptr};
  double yy[MN], ff[MN], JJ[MN * MN];
  auto eval_at = [&](long long i, const double* add /*SoA offset or null*/, double scalar, double* out) {
    for (int e = 0; e < MN; e++) {
      double v = S.y[(size_t)e * M + i];
      if (add) v = v + add[(size_t)e * M + i];
      else if (scalar != 0.0) v = v + scalar;
      yy[e] = v;
    }
    dyns(5, par, sign, MN, 1, yy, ff);
    for (int e = 0; e < MN; e++) out[(size_t)e * M + i] = ff[e];
  };
  for (long long i = 0; i < M; i++) {
    ph_init(S, P, i, y0);
    eval_at(i, nullptr, 0.0, S.f0);
    ph_init_b(S, i);
  }
  for (;;) {
    long long active = 0;
    for (long long i = 0; i < M; i++) {
      ph_head(S, P, i);
      if (sc[i].phase != PH_FINISHED) active++;
      if (sc[i].phase == PH_WANT_JAC) {
        for (int e = 0; e < MN; e++) yy[e] = S.y[(size_t)e * M + i];
        jacf(5, par, sign, MN, 1, yy, JJ);
        for (int k = 0; k < MN * MN; k++) S.Jac[(size_t)i * MN * MN + k] = JJ[k];
      }
    }
    if (!active) break;
    for (long long i = 0; i < M; i++) ph_prep(S, P, i, (i & 1) != 0);   // odd rows: deferred factorisation (GPU protocol)
    for (long long i = 0; i < M; i++) ph_lu(S, i);
    for (;;) {
      long long n = 0;
      for (long long i = 0; i < M; i++)
        if (sc[i].phase == PH_NEWTON) {
          n++;
          for (int q = 0; q < sc[i].NbrStg; q++) eval_at(i, S.z + (size_t)q * MN * M, 0.0, S.F + (size_t)q * MN * M);
        }
      if (!n) break;
      for (long long i = 0; i < M; i++) ph_newton(S, P, i);
    }
    for (long long i = 0; i < M; i++)
      if (sc[i].phase == PH_ERR_EVAL) {
        eval_at(i, nullptr, S.errs[i], S.ev);
        ph_estrad_b(S, P, i);
      }
    for (long long i = 0; i < M; i++)
      if (sc[i].phase == PH_F0_EVAL) {
        eval_at(i, nullptr, 0.0, S.f0);
        ph_accept_b(S, P, i);
      }
  }
  for (long long i = 0; i < M; i++) {
    for (int e = 0; e < MN; e++) y_out[i * MN + e] = S.y[(size_t)e * M + i];
    t_out[i] = sc[i].t; h_out[i] = sc[i].h; res[i] = sc[i].res; nstg[i] = sc[i].NbrStg;
    for (int k = 0; k < 8; k++) stats[i * 8 + k] = sc[i].st[k];
    if (work) work[i] = sc[i].work;
  }
  return 0;
}
