// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Driver around the UNMODIFIED reference solver: compiles /root/reference/src/cpp/radaute.hpp +
// radaute_impl.hpp (HEADER_ONLY, as examples/cpp/radaute_hamiltonian_example.cpp:1-5 does) and
// janus_ode_common.hpp where they lie, against the image's libtorch and the stand-in headers of
// oracle/ref_shim/ (the un-vendored janus / Eigen / SUNDIALS / PETSc includes).  No reference source is
// copied: the headers are #included from /root/reference by the Makefile's -I.
//
// What it runs: janus::RadauTe(OdeFcn, JacFn, tspan, y0, Op, params).solve()
// (src/cpp/radaute.hpp:424-427) ONE TRAJECTORY PER SOLVER OBJECT (M = 1), because RadauTe's batch-coupled
// `if (m1_12.all()==false) continue;` (radaute_impl.hpp:1006) stalls any batch whose
// trajectories desynchronise; with M = 1 the masked code is the per-trajectory algorithm.
// `batch=1` runs the whole input as one RadauTe batch instead (to demonstrate Q3).
//
// OdeFcn / JacFn are either
//   adaptor 1: the reference's own evalDyns / evalJac (+U twins) (janus_ode_common.hpp:2161-2299) on a
//              torch Hamiltonian written like the reference's fixtures (tests/cpp/ode_solver_test.cpp:170-197,
//              examples/cpp/radaute_hamiltonian_example.cpp:28-39) -- i.e. the full reference hot path; or
//   adaptor 0: closed-form dynamics / Jacobian (the form of the reference's vdpdyns / vdpjac fixtures,
//              tests/cpp/ode_solver_test.cpp:204-334) in the operation order of oracle/problems.py.
#define HEADER_ONLY
#include "radaute.hpp"
#include <janus/tensordual.hpp>
#include "janus_ode_common.hpp"

#include <cstdint>
#include <cstring>
#include <csignal>
#include <fcntl.h>
#include <unistd.h>

using torch::Tensor;
using Slice = torch::indexing::Slice;

namespace {

enum { HAM_HARMONIC = 0, HAM_VDP = 1, HAM_PMP_SMOOTH = 2, HAM_PMP_BANG = 3 };

struct Problem {
  int ham;
  double par;      // omega (harmonic), mu (vdP), W (PMP)
  int sign;        // +1 reference, -1 canonical (realised through the sign of the x-part of H)
  int adaptor;
};

// ---- torch Hamiltonians (adaptor 1).  evalDyns hard-codes pdot = +dH/dx, so the canonical sign is
// obtained from H~(x,p) with dH~/dx = -dH/dx, dH~/dp = dH/dp: only separable H needs it here (harmonic).
Tensor H_harmonic(const Tensor& x, const Tensor& p, double w, int sign) {
  auto p1 = p.index({Slice(), 0});
  auto x1 = x.index({Slice(), 0});
  if (sign > 0) return (p1.pow(2) + (w * w) * x1.pow(2)) / 2;
  return (p1.pow(2) - (w * w) * x1.pow(2)) / 2;
}
Tensor H_vdp(const Tensor& x, const Tensor& p, double mu) {
  auto p1 = p.index({Slice(), 0}), p2 = p.index({Slice(), 1});
  auto x1 = x.index({Slice(), 0}), x2 = x.index({Slice(), 1});
  return p1 * x2 + p2 * (mu * (1 - x1 * x1) * x2 - x1);
}
Tensor Hu_pmp(const Tensor& x, const Tensor& p, const Tensor& u, double W) {
  auto p1 = p.index({Slice(), 0}), p2 = p.index({Slice(), 1});
  auto x1 = x.index({Slice(), 0}), x2 = x.index({Slice(), 1});
  return p1 * x2 + p2 * u * ((1 - x1 * x1) * x2 - x1) + (W / 2) * u * u + 1;
}
Tensor control(const Tensor& y, int ham, double W) {
  torch::NoGradGuard ng;
  auto p2 = y.index({Slice(), 1}), x1 = y.index({Slice(), 2}), x2 = y.index({Slice(), 3});
  if (ham == HAM_PMP_SMOOTH) return (-p2 * ((1 - x1 * x1) * x2 - x1) / W).detach();
  auto m = p2 * (1 - x1 * x1) * x2 < 0;
  return torch::where(m, torch::full_like(p2, 100.0), torch::full_like(p2, 1.0)).detach();
}

// ---- closed forms (adaptor 0), operation order of oracle/problems.py
Tensor closed_dyns(const Problem& P, const Tensor& y) {
  torch::NoGradGuard ng;
  const double s = P.sign;
  if (P.ham == HAM_HARMONIC) {
    auto p = y.index({Slice(), 0}), x = y.index({Slice(), 1});
    return torch::stack({(s * P.par * P.par) * x, p}, 1);
  }
  auto p1 = y.index({Slice(), 0}), p2 = y.index({Slice(), 1});
  auto x1 = y.index({Slice(), 2}), x2 = y.index({Slice(), 3});
  if (P.ham == HAM_VDP) {
    const double mu = P.par;
    return torch::stack({s * (p2 * (-2.0 * mu * x1 * x2 - 1.0)),
                         s * (p1 + p2 * (mu * (1.0 - x1 * x1))),
                         x2.clone(),
                         mu * (1.0 - x1 * x1) * x2 - x1}, 1);
  }
  auto u = control(y, P.ham, P.par);
  return torch::stack({s * (p2 * u * (-2.0 * x1 * x2 - 1.0)),
                       s * (p1 + p2 * u * (1.0 - x1 * x1)),
                       x2.clone(),
                       u * ((1.0 - x1 * x1) * x2 - x1)}, 1);
}
Tensor closed_jac(const Problem& P, const Tensor& y) {
  torch::NoGradGuard ng;
  const double s = P.sign;
  const int64_t M = y.size(0);
  if (P.ham == HAM_HARMONIC) {
    auto J = torch::zeros({M, 2, 2}, torch::kFloat64);
    J.index_put_({Slice(), 0, 1}, s * P.par * P.par);
    J.index_put_({Slice(), 1, 0}, 1.0);
    return J;
  }
  auto J = torch::zeros({M, 4, 4}, torch::kFloat64);
  auto p2 = y.index({Slice(), 1});
  auto x1 = y.index({Slice(), 2}), x2 = y.index({Slice(), 3});
  Tensor g;   // mu (vdP) or u (PMP), multiplying the (1-x1^2) x2 terms
  if (P.ham == HAM_VDP) {
    const double mu = P.par;
    J.index_put_({Slice(), 0, 1}, s * (-2.0 * mu * x1 * x2 - 1.0));
    J.index_put_({Slice(), 0, 2}, s * (p2 * (-2.0 * mu * x2)));
    J.index_put_({Slice(), 0, 3}, s * (p2 * (-2.0 * mu * x1)));
    J.index_put_({Slice(), 1, 0}, s * 1.0);
    J.index_put_({Slice(), 1, 1}, s * (mu * (1.0 - x1 * x1)));
    J.index_put_({Slice(), 1, 2}, s * (p2 * mu * (-2.0 * x1)));
    J.index_put_({Slice(), 2, 3}, 1.0);
    J.index_put_({Slice(), 3, 2}, -2.0 * mu * x1 * x2 - 1.0);
    J.index_put_({Slice(), 3, 3}, mu * (1.0 - x1 * x1));
    return J;
  }
  auto u = control(y, P.ham, P.par);
  J.index_put_({Slice(), 0, 1}, s * (u * (-2.0 * x1 * x2 - 1.0)));
  J.index_put_({Slice(), 0, 2}, s * (p2 * u * (-2.0 * x2)));
  J.index_put_({Slice(), 0, 3}, s * (p2 * u * (-2.0 * x1)));
  J.index_put_({Slice(), 1, 0}, s * 1.0);
  J.index_put_({Slice(), 1, 1}, s * (u * (1.0 - x1 * x1)));
  J.index_put_({Slice(), 1, 2}, s * (p2 * u * (-2.0 * x1)));
  J.index_put_({Slice(), 2, 3}, 1.0);
  J.index_put_({Slice(), 3, 2}, u * (-2.0 * x1 * x2 - 1.0));
  J.index_put_({Slice(), 3, 3}, u * (1.0 - x1 * x1));
  return J;
}

Tensor adaptor_dyns(const Problem& P, const Tensor& y) {
  using HFn = std::function<Tensor(const Tensor&, const Tensor&, double)>;
  using HuFn = std::function<Tensor(const Tensor&, const Tensor&, const Tensor&, double)>;
  if (P.ham == HAM_HARMONIC) {
    const int sign = P.sign;
    HFn H = [sign](const Tensor& x, const Tensor& p, double w) { return H_harmonic(x, p, w, sign); };
    return janus::evalDyns<double>(y, P.par, H).detach();
  }
  if (P.ham == HAM_VDP) {
    HFn H = [](const Tensor& x, const Tensor& p, double mu) { return H_vdp(x, p, mu); };
    return janus::evalDyns<double>(y, P.par, H).detach();
  }
  HuFn Hu = [](const Tensor& x, const Tensor& p, const Tensor& u, double W) { return Hu_pmp(x, p, u, W); };
  return janus::evalDynsU<double>(y, control(y, P.ham, P.par), P.par, Hu).detach();
}
Tensor adaptor_jac(const Problem& P, const Tensor& y) {
  using HFn = std::function<Tensor(const Tensor&, const Tensor&, double)>;
  using HuFn = std::function<Tensor(const Tensor&, const Tensor&, const Tensor&, double)>;
  if (P.ham == HAM_HARMONIC) {
    const int sign = P.sign;
    HFn H = [sign](const Tensor& x, const Tensor& p, double w) { return H_harmonic(x, p, w, sign); };
    return janus::evalJac<double>(y, P.par, H).detach();
  }
  if (P.ham == HAM_VDP) {
    HFn H = [](const Tensor& x, const Tensor& p, double mu) { return H_vdp(x, p, mu); };
    return janus::evalJac<double>(y, P.par, H).detach();
  }
  HuFn Hu = [](const Tensor& x, const Tensor& p, const Tensor& u, double W) { return Hu_pmp(x, p, u, W); };
  return janus::evalJacU<double>(y, control(y, P.ham, P.par), P.par, Hu).detach();
}

}  // namespace

static std::string g_err;

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// Evaluate OdeFcn / JacFn alone (pins the Hamiltonian adaptor): ydot [M,Ny], jac [M,Ny,Ny].
int ref_eval(int ham, double par, int sign, int adaptor, int64_t M, int Ny, const double* y,
             double* ydot, double* jac) {
  Problem P{ham, par, sign, adaptor};
  if (adaptor && sign < 0 && ham != HAM_HARMONIC) return -2;
  auto yt = torch::from_blob((void*)y, {M, Ny}, torch::kFloat64).clone();
  Tensor f = adaptor ? adaptor_dyns(P, yt) : closed_dyns(P, yt);
  Tensor J = adaptor ? adaptor_jac(P, yt) : closed_jac(P, yt);
  std::memcpy(ydot, f.contiguous().data_ptr<double>(), sizeof(double) * M * Ny);
  std::memcpy(jac, J.contiguous().data_ptr<double>(), sizeof(double) * M * Ny * Ny);
  return 0;
}

// Solve rows [0,M) of (y0 [M,Ny], tspan [M,2]).  Outputs per row: y [Ny], t, h, res, stats[8] in the
// order of StatsTe (radaute.hpp:98-108: FcnNbr, JacNbr, DecompNbr, SolveNbr, StepNbr, AccptNbr,
// StepRejNbr, NewtRejNbr), final NbrStg, and the solver's pass counter `count`.
int ref_radau_solve(int ham, double par, int sign, int adaptor, int batch,
                    int64_t M, int Ny, const double* y0, const double* tspan,
                    double rtol, double atol, double h0, int nbrstg, int minstg, int maxstg,
                    int max_steps,
                    double* y, double* t, double* h, double* res, int64_t* stats, int64_t* stg,
                    int64_t* passes) {
  Problem P{ham, par, sign, adaptor};
  if (adaptor && sign < 0 && ham != HAM_HARMONIC) return -2;
  torch::set_num_threads(1);
  // the reference dumps tensors to std::cerr on several paths (radaute_impl.hpp:1837-1849 etc.)
  // and to std::cout / printf (:1006-1020 ...): silence both streams at the descriptor level for the call
  std::cerr.setstate(std::ios_base::failbit);
  std::cout.flush(); fflush(stdout); fflush(stderr);
  const int fd_out = dup(1), fd_err = dup(2), fd_null = open("/dev/null", O_WRONLY);
  dup2(fd_null, 1); dup2(fd_null, 2);
  janus::OdeFnType f = [P](const Tensor&, const Tensor& yy, const Tensor&) {
    return P.adaptor ? adaptor_dyns(P, yy) : closed_dyns(P, yy);
  };
  janus::JacFnType jf = [P](const Tensor&, const Tensor& yy, const Tensor&) {
    return P.adaptor ? adaptor_jac(P, yy) : closed_jac(P, yy);
  };
  const int64_t nsolve = batch ? 1 : M;
  const int64_t rows = batch ? M : 1;
  int rc = 0;
  try {
  for (int64_t i = 0; i < nsolve; ++i) {
    auto yt = torch::from_blob((void*)(y0 + i * Ny), {rows, Ny}, torch::kFloat64).clone();
    auto ts = torch::from_blob((void*)(tspan + i * 2), {rows, 2}, torch::kFloat64).clone();
    janus::OptionsTe op;
    op.RelTol = torch::tensor({rtol}, torch::kFloat64);
    op.AbsTol = torch::tensor({atol}, torch::kFloat64);
    op.InitialStep = torch::tensor({h0}, torch::kFloat64);
    op.NbrStg = torch::tensor(nbrstg, torch::kInt64);
    op.MinNbrStg = torch::tensor(minstg, torch::kInt64);
    op.MaxNbrStg = torch::tensor(maxstg, torch::kInt64);
    op.MaxNbrStep = max_steps;
    auto params = torch::empty({0}, torch::kFloat64);
    janus::RadauTe r(f, jf, ts, yt, op, params);
    Tensor rs = r.solve();
    auto put = [&](double* dst, const Tensor& src, int64_t n) {
      auto c = src.to(torch::kFloat64).contiguous();
      std::memcpy(dst, c.data_ptr<double>(), sizeof(double) * n);
    };
    put(y + i * Ny, r.y, rows * Ny);
    put(t + i, r.t, rows);
    put(h + i, r.h, rows);
    put(res + i, rs, rows);
    const Tensor* st[8] = {&r.stats.FcnNbr, &r.stats.JacNbr, &r.stats.DecompNbr, &r.stats.SolveNbr,
                           &r.stats.StepNbr, &r.stats.AccptNbr, &r.stats.StepRejNbr, &r.stats.NewtRejNbr};
    for (int64_t k = 0; k < rows; ++k) {
      for (int c = 0; c < 8; ++c)
        stats[(i + k) * 8 + c] = st[c]->defined() ? st[c]->to(torch::kInt64).reshape({-1})[k].item<int64_t>() : -1;
      stg[i + k] = r.NbrStg.to(torch::kInt64).reshape({-1})[k].item<int64_t>();
      passes[i + k] = r.count;
    }
  }
  } catch (const std::exception& e) { g_err = e.what(); rc = -3; }
  std::cout.flush(); fflush(stdout); fflush(stderr);
  dup2(fd_out, 1); dup2(fd_err, 2); close(fd_out); close(fd_err); close(fd_null);
  std::cerr.clear();
  return rc;
}

// One trajectory with a log of every OdeFcn call: row = [pass counter `count`, t argument, r.t, r.h,
// NbrStg, Newt, Theta, err, FacConv, hnew] read from the solver object's public fields at call time
// (radaute.hpp:244-420).  Used to find the pass at which a restatement leaves the reference.
int ref_radau_trace(int ham, double par, int sign, int adaptor, int Ny, const double* y0, const double* tspan,
                    double rtol, double atol, double h0, int nbrstg, int minstg, int maxstg, int max_steps,
                    double* log, int cap, int* nlog, double* y, int64_t* stats) {
  Problem P{ham, par, sign, adaptor};
  torch::set_num_threads(1);
  std::cerr.setstate(std::ios_base::failbit);
  std::cout.flush(); fflush(stdout); fflush(stderr);
  const int fd_out = dup(1), fd_err = dup(2), fd_null = open("/dev/null", O_WRONLY);
  dup2(fd_null, 1); dup2(fd_null, 2);
  janus::RadauTe* rp = nullptr;
  int n = 0;
  auto val = [](const Tensor& a) { return a.defined() && a.numel() >= 1 ? a.reshape({-1})[0].item<double>() : NAN; };
  janus::OdeFnType f = [&](const Tensor& tt, const Tensor& yy, const Tensor&) {
    if (rp && n < cap) {
      double* L = log + 10 * n++;
      L[0] = rp->count; L[1] = val(tt); L[2] = val(rp->t); L[3] = val(rp->h); L[4] = val(rp->NbrStg);
      L[5] = val(rp->Newt); L[6] = val(rp->Theta); L[7] = val(rp->err); L[8] = val(rp->FacConv); L[9] = val(rp->hnew);
    }
    return P.adaptor ? adaptor_dyns(P, yy) : closed_dyns(P, yy);
  };
  janus::JacFnType jf = [P](const Tensor&, const Tensor& yy, const Tensor&) {
    return P.adaptor ? adaptor_jac(P, yy) : closed_jac(P, yy);
  };
  auto yt = torch::from_blob((void*)y0, {1, Ny}, torch::kFloat64).clone();
  auto ts = torch::from_blob((void*)tspan, {1, 2}, torch::kFloat64).clone();
  janus::OptionsTe op;
  op.RelTol = torch::tensor({rtol}, torch::kFloat64);
  op.AbsTol = torch::tensor({atol}, torch::kFloat64);
  op.InitialStep = torch::tensor({h0}, torch::kFloat64);
  op.NbrStg = torch::tensor(nbrstg, torch::kInt64);
  op.MinNbrStg = torch::tensor(minstg, torch::kInt64);
  op.MaxNbrStg = torch::tensor(maxstg, torch::kInt64);
  op.MaxNbrStep = max_steps;
  auto params = torch::empty({0}, torch::kFloat64);
  janus::RadauTe r(f, jf, ts, yt, op, params);
  rp = &r;
  r.solve();
  std::memcpy(y, r.y.contiguous().data_ptr<double>(), sizeof(double) * Ny);
  const Tensor* st[8] = {&r.stats.FcnNbr, &r.stats.JacNbr, &r.stats.DecompNbr, &r.stats.SolveNbr,
                         &r.stats.StepNbr, &r.stats.AccptNbr, &r.stats.StepRejNbr, &r.stats.NewtRejNbr};
  for (int c = 0; c < 8; ++c) stats[c] = st[c]->defined() ? st[c]->to(torch::kInt64).reshape({-1})[0].item<int64_t>() : -1;
  *nlog = n;
  std::cout.flush(); fflush(stdout); fflush(stderr);
  dup2(fd_out, 1); dup2(fd_err, 2); close(fd_out); close(fd_err); close(fd_null);
  std::cerr.clear();
  return 0;
}

// One trajectory with the reference's own record of the computed points: RadauTe's OutFlag 1 buffers tout [1,nout],
// yout [1,nout,Ny], nout (radaute_impl.hpp:288-296, 1166-1183), copied out up to `cap` points.
int ref_radau_steps(int ham, double par, int sign, int adaptor, int Ny, const double* y0, const double* tspan,
                    double rtol, double atol, double h0, int nbrstg, int minstg, int maxstg, int max_steps,
                    int cap, double* tout, double* yout, int* nout, double* y, int64_t* stats) {
  Problem P{ham, par, sign, adaptor};
  torch::set_num_threads(1);
  std::cerr.setstate(std::ios_base::failbit);
  std::cout.flush(); fflush(stdout); fflush(stderr);
  const int fd_out = dup(1), fd_err = dup(2), fd_null = open("/dev/null", O_WRONLY);
  dup2(fd_null, 1); dup2(fd_null, 2);
  int rc = 0;
  try {
    janus::OdeFnType f = [P](const Tensor&, const Tensor& yy, const Tensor&) {
      return P.adaptor ? adaptor_dyns(P, yy) : closed_dyns(P, yy);
    };
    janus::JacFnType jf = [P](const Tensor&, const Tensor& yy, const Tensor&) {
      return P.adaptor ? adaptor_jac(P, yy) : closed_jac(P, yy);
    };
    auto yt = torch::from_blob((void*)y0, {1, Ny}, torch::kFloat64).clone();
    auto ts = torch::from_blob((void*)tspan, {1, 2}, torch::kFloat64).clone();
    janus::OptionsTe op;
    op.RelTol = torch::tensor({rtol}, torch::kFloat64);
    op.AbsTol = torch::tensor({atol}, torch::kFloat64);
    op.InitialStep = torch::tensor({h0}, torch::kFloat64);
    op.NbrStg = torch::tensor(nbrstg, torch::kInt64);
    op.MinNbrStg = torch::tensor(minstg, torch::kInt64);
    op.MaxNbrStg = torch::tensor(maxstg, torch::kInt64);
    op.MaxNbrStep = max_steps;
    auto params = torch::empty({0}, torch::kFloat64);
    janus::RadauTe r(f, jf, ts, yt, op, params);
    r.solve();
    const int n = (int)r.nout.reshape({-1})[0].item<int64_t>();
    *nout = n;
    auto to = r.tout.to(torch::kFloat64).contiguous();
    auto yo = r.yout.to(torch::kFloat64).contiguous();
    for (int k = 0; k < n && k < cap; ++k) {
      tout[k] = to[0][k].item<double>();
      for (int i = 0; i < Ny; ++i) yout[k * Ny + i] = yo[0][k][i].item<double>();
    }
    std::memcpy(y, r.y.contiguous().data_ptr<double>(), sizeof(double) * Ny);
    const Tensor* st[8] = {&r.stats.FcnNbr, &r.stats.JacNbr, &r.stats.DecompNbr, &r.stats.SolveNbr,
                           &r.stats.StepNbr, &r.stats.AccptNbr, &r.stats.StepRejNbr, &r.stats.NewtRejNbr};
    for (int c = 0; c < 8; ++c) stats[c] = st[c]->defined() ? st[c]->to(torch::kInt64).reshape({-1})[0].item<int64_t>() : -1;
  } catch (const std::exception& e) { g_err = e.what(); rc = -3; }
  std::cout.flush(); fflush(stdout); fflush(stderr);
  dup2(fd_out, 1); dup2(fd_err, 2); close(fd_out); close(fd_err); close(fd_null);
  std::cerr.clear();
  return rc;
}

// evalDynsDual / evalJacDual (+ U twins) of the reference (janus_ode_common.hpp:2075-2158, :2302-2416) on the fixture
// Hamiltonians: y [M,Ny], yd [M,Ny,D] -> f [M,Ny], fd [M,Ny,D], J [M,Ny,Ny], Jd [M,Ny,Ny,D].  These run the eight
// third-order autograd ladders (:787-2072).
int ref_eval_dual(int ham, double par, int64_t M, int Ny, int D, const double* y, const double* yd, double* f, double* fd,
                  double* J, double* Jd) {
  using HFn = std::function<Tensor(const Tensor&, const Tensor&, double)>;
  using HuFn = std::function<Tensor(const Tensor&, const Tensor&, const Tensor&, double)>;
  int rc = 0;
  try {
    auto yt = torch::from_blob((void*)y, {M, Ny}, torch::kFloat64).clone();
    auto ydt = torch::from_blob((void*)yd, {M, Ny, D}, torch::kFloat64).clone();
    janus::TensorDual yy(yt, ydt);
    janus::TensorDual fr;
    janus::TensorMatDual Jr;
    if (ham == HAM_HARMONIC) {
      HFn H = [](const Tensor& x, const Tensor& p, double w) { return H_harmonic(x, p, w, +1); };
      fr = janus::evalDynsDual<double>(yy, par, H);
      Jr = janus::evalJacDual<double>(yy, par, H);
    } else if (ham == HAM_VDP) {
      HFn H = [](const Tensor& x, const Tensor& p, double mu) { return H_vdp(x, p, mu); };
      fr = janus::evalDynsDual<double>(yy, par, H);
      Jr = janus::evalJacDual<double>(yy, par, H);
    } else {
      HuFn Hu = [](const Tensor& x, const Tensor& p, const Tensor& u, double W) { return Hu_pmp(x, p, u, W); };
      auto u = control(yt, ham, par);
      fr = janus::evalDynsUDual<double>(yy, u, par, Hu);
      Jr = janus::evalJacDualU<double>(yy, u, par, Hu);
    }
    auto cp = [](double* dst, const Tensor& src) {
      auto c = src.detach().to(torch::kFloat64).contiguous();
      std::memcpy(dst, c.data_ptr<double>(), sizeof(double) * c.numel());
    };
    cp(f, fr.r); cp(fd, fr.d); cp(J, Jr.r); cp(Jd, Jr.d);
  } catch (const std::exception& e) { g_err = e.what(); rc = -3; }
  return rc;
}

}  // extern "C"
