/*
 * Scalar C restatement of the reference's RADAU stepper  --  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under janus_ode_b200/ links, loads or calls this file.  It is the parity checker
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: "parity unpinned" for final states / accepted-step counts: the reference
 * (optstat/janus-ode) cannot be built here (un-vendored janus-tensor-dual / janus-linear,
 * Eigen3, SUNDIALS, PETSc) and ships no golden vectors of a solve.  Pinned instead: the
 * coefficient tables (oracle/tools/gen_tables.py), the reference's closed-form GTest
 * fixtures, SciPy Radau at tolerance level, and agreement with the independent Python
 * restatement oracle/radau_scalar.py (tests/test_oracle_*.py).
 *
 * Restates, per trajectory:
 *   control flow   /root/reference/src/matlab/radau.m:630-1256  (== masked form in
 *                  src/cpp/radauted_impl.hpp:620-1645, src/cpp/radaute_impl.hpp:537-1345)
 *   DecomRC        src/cpp/radaute_impl.hpp:1502-1661   LU with partial pivoting, LAPACK
 *                  zgetf2 rules: pivot = first max |re|+|im| (izamax), column scaled by the
 *                  reciprocal of the pivot; the real block is done in real arithmetic (the
 *                  reference promotes it to complex128 with zero imaginary part: same pivots,
 *                  same values)
 *   Solvrad        src/cpp/radaute_impl.hpp:1721-1897
 *   Estrad         src/cpp/radaute_impl.hpp:1902-1986
 *   tables         src/cpp/radaute_impl.hpp:2187-2644  -> radau_tables.h (generated)
 *   Hamiltonian dynamics: hand-derived closed forms of ydot = [sigma dH/dx ; dH/dp] and
 *                  J = d ydot/d y  (src/cpp/janus_ode_common.hpp:2161-2299); fixtures
 *                  tests/cpp/ode_solver_test.cpp:228-248,299-334
 * Deviations Q3,Q4,Q6,Q7,Q8,Q9: see oracle/radau_scalar.py header (identical here).
 *
 * Compile with -ffp-contract=off: the reference's libtorch CPU ops never fuse mul+add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "radau_tables.h"

#define NYMAX 16
#define SMAX 7
#define EPS 2.220446049250313e-16

/*
 * Arithmetic modes (same control flow, different roundoff realisation):
 *   g_mirror = 0  "reference": what the reference's libtorch CPU ops do -- no fused multiply-add,
 *                 true divisions, libm pow, closed-form derivatives.
 *   g_mirror = 1  "mirror": the documented ARITHMETIC CONTRACT of the CUDA path
 *                 (janus_ode_b200/csrc/{dual,lu,radau_core,detmath}.cuh): fma() in the same
 *                 places, reciprocal-multiplies, deterministic pow/log/exp built from IEEE basic
 *                 ops, derivatives of H by (dense) dual / hyper-dual arithmetic.  Bit-identical
 *                 to the GPU.  Needed because the algorithm's accepted-step counts are sensitive
 *                 to the last bit (tests/test_oracle_sensitivity.py: recompiling THIS file with
 *                 FMA contraction alone changes >1% of the counts).
 * Set once before a (parallel) solve; read-only inside.
 */
static int g_mirror = 0;
#define FMA(a, b, c) (g_mirror ? fma((a), (b), (c)) : ((a) * (b) + (c)))

/* ---- deterministic log/exp/pow: same sequence as csrc/detmath.cuh ---- */
static const double DET_EXP2_TAB[32] = {   /* 2^(j/32), correctly rounded: csrc/detmath.cuh JB_EXP2_TAB_VALUES */
  0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
  0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
  0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
  0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
  0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
  0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
  0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
  0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
static const double DET_LOG_INVC[32] = {   /* csrc/detmath.cuh JB_LOG_INVC_VALUES */
  0x1.f81f81f81f820p-1, 0x1.e9131abf0b767p-1, 0x1.dae6076b981dbp-1, 0x1.cd85689039b0bp-1,
  0x1.c0e070381c0e0p-1, 0x1.b4e81b4e81b4fp-1, 0x1.a98ef606a63bep-1, 0x1.9ec8e951033d9p-1,
  0x1.948b0fcd6e9e0p-1, 0x1.8acb90f6bf3aap-1, 0x1.8181818181818p-1, 0x1.78a4c8178a4c8p-1,
  0x1.702e05c0b8170p-1, 0x1.6816816816817p-1, 0x1.6058160581606p-1, 0x1.58ed2308158edp-1,
  0x1.51d07eae2f815p-1, 0x1.4afd6a052bf5bp-1, 0x1.446f86562d9fbp-1, 0x1.3e22cbce4a902p-1,
  0x1.3813813813814p-1, 0x1.323e34a2b10bfp-1, 0x1.2c9fb4d812ca0p-1, 0x1.27350b8812735p-1,
  0x1.21fb78121fb78p-1, 0x1.1cf06ada2811dp-1, 0x1.1811811811812p-1, 0x1.135c81135c811p-1,
  0x1.0ecf56be69c90p-1, 0x1.0a6810a6810a7p-1, 0x1.0624dd2f1a9fcp-1, 0x1.0204081020408p-1};
static const double DET_LOG_LOGC[32] = {   /* csrc/detmath.cuh JB_LOG_LOGC_VALUES */
  0x1.fc0a8b0fc03c4p-7, 0x1.77458f632dcffp-5, 0x1.341d7961bd1d0p-4, 0x1.a926d3a4ad562p-4,
  0x1.0d77e7cd08e5bp-3, 0x1.44d2b6ccb7d1cp-3, 0x1.7ab890210d907p-3, 0x1.af3c94e80bff3p-3,
  0x1.e27076e2af2e8p-3, 0x1.0a324e27390e2p-2, 0x1.22941fbcf7966p-2, 0x1.3a64c556945eap-2,
  0x1.51aad872df82ep-2, 0x1.686c81e9b14adp-2, 0x1.7eaf83b82afc2p-2, 0x1.947941c2116fbp-2,
  0x1.a9cec9a9a084ap-2, 0x1.beb4d9da71b7ap-2, 0x1.d32fe7e00ebd5p-2, 0x1.e744261d68789p-2,
  0x1.faf588f78f31dp-2, 0x1.0723e5c1cdf41p-1, 0x1.109f39e2d4c96p-1, 0x1.19ee6b467c96fp-1,
  0x1.23130d7bebf43p-1, 0x1.2c0e9ed448e8cp-1, 0x1.34e289d9ce1d2p-1, 0x1.3d9026a7156fbp-1,
  0x1.4618bc21c5ec2p-1, 0x1.4e7d811b75bb0p-1, 0x1.56bf9d5b3f399p-1, 0x1.5ee02a9241676p-1};
static double det_log(double x) {   /* csrc/detmath.cuh det_log, operation by operation */
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  int64_t b;
  memcpy(&b, &x, 8);
  const int e = (int)((b >> 52) & 0x7ff) - 1023;
  const int j = (int)((b >> 47) & 31);
  int64_t mb = (b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
  double m;
  memcpy(&m, &mb, 8);
  const double r = fma(m, DET_LOG_INVC[j], -1.0);
  double p = -1.0 / 8.0;
  p = fma(p, r, 1.0 / 7.0);
  p = fma(p, r, -1.0 / 6.0);
  p = fma(p, r, 1.0 / 5.0);
  p = fma(p, r, -1.0 / 4.0);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  const double ed = (double)e;
  const double hi = fma(ed, LN2_HI, DET_LOG_LOGC[j]);
  const double lo = fma(r * r, p, fma(ed, LN2_LO, r));
  return hi + lo;
}
static double det_exp(double t) {   /* csrc/detmath.cuh det_exp, operation by operation */
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double SHIFT = 6755399441055744.0;
  const double ns = fma(t, 4.6166241308446828e+01, SHIFT);
  const double nd = ns - SHIFT;
  double r = fma(nd, -(LN2_HI / 32.0), t);
  r = fma(nd, -(LN2_LO / 32.0), r);
  int64_t nb;
  memcpy(&nb, &ns, 8);
  const int n = (int)(uint32_t)(nb & 0xffffffffLL);
  const int j = n & 31, k = n >> 5;
  double q = 1.0 / 720.0;
  q = fma(q, r, 1.0 / 120.0);
  q = fma(q, r, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  q = q * DET_EXP2_TAB[j];
  int64_t qb;
  memcpy(&qb, &q, 8);
  qb += ((int64_t)k << 52);
  double out;
  memcpy(&out, &qb, 8);
  return out;
}
static double det_pow(double x, double y) { return det_exp(y * det_log(x)); }
static double det_tanh(double x) {   /* csrc/detmath.cuh det_tanh_t, operation by operation */
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double SHIFT = 6755399441055744.0;
  double ax = fabs(x);
  ax = ax > 20.0 ? 20.0 : ax;
  const double ns = fma(ax, -9.2332482616893657e+01, SHIFT);
  const double nd = ns - SHIFT;
  double r = fma(nd, -(LN2_HI / 32.0), -2.0 * ax);
  r = fma(nd, -(LN2_LO / 32.0), r);
  int64_t nb;
  memcpy(&nb, &ns, 8);
  const int n = (int)(uint32_t)(nb & 0xffffffffLL);
  const int j = n & 31, k = n >> 5;
  double p = 1.0 / 720.0;
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double q = p * DET_EXP2_TAB[j];
  int64_t qb;
  memcpy(&qb, &q, 8);
  qb += ((int64_t)k << 52);
  double e;
  memcpy(&e, &qb, 8);
  const double num = 1.0 - e, d = 1.0 + e;
  double y = fma(fma(32.0 / 99.0, d, -48.0 / 33.0), d, 70.0 / 33.0);
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  double th = num * y;
  th = fma(fma(-d, th, num), y, th);
  th = x < 0 ? -th : th;
  return x != x ? x : th;
}
static double POW(double x, double y) { return g_mirror ? det_pow(x, y) : pow(x, y); }

/* ---- problem ids (shared numbering with include/janus_b200.h JB200_HAM_*) ---- */
enum {
  PB_HARMONIC = 0,      /* par[0]=omega                      N=1 */
  PB_VDP = 1,           /* par[0]=mu                         N=2 */
  PB_VDP_PMP_SMOOTH = 2,/* par[0]=W          Hu test form, u=-p2((1-x1^2)x2-x1)/W frozen */
  PB_VDP_PMP_BANG = 3,  /* par[0]=W,par[1]=umin,par[2]=umax  Hu test form, bang-bang u   */
  PB_VDP_AUG_BANG = 4,  /* par[0]=umin,par[1]=umax  H=p1x2+p2(u(1-x1^2)x2-x1), bang-bang */
  PB_MLP = 5            /* par = [N, width, W1,b1,W2,b2,W3,b3,W4,b4]  tanh MLP, input [x;p] */
};

typedef struct {
  double RelTol[NYMAX], AbsTol[NYMAX];
  double InitialStep, JacRecompute;
  int Start_Newt, MaxNbrNewton, NbrStg, MinNbrStg, MaxNbrStg;
  long long MaxNbrStep;
  double Safe, Quot1, Quot2, FacL, FacR, Vitu, Vitd, hhou, hhod;
  int Gustafsson;
} oracle_opts;

typedef struct {
  int id, Ny;
  double sign;
  const double *par;
} problem_t;

/* ------------------------------------------------------------------ dynamics */
static double bang(double p2, double x1, double x2, double umin, double umax) {
  return (p2 * (1.0 - x1 * x1) * x2 < 0.0) ? umax : umin;
}

static void mlp_eval(const problem_t *pb, const double *y, double *f, double *J);
static void mlp_eval_mirror(const problem_t *pb, const double *y, double *f, double *J);

/* ------------------------------------------------------------------ mirror-mode derivatives
 * Dense dual / hyper-dual arithmetic following the ARITHMETIC CONTRACT of csrc/dual.cuh (the
 * GPU's compile-time-sparse numbers skip structurally-zero entries; adding / multiplying exact
 * zeros here gives the same bits).  The Hamiltonians below repeat the expression trees of
 * csrc/hamiltonians.cuh operation by operation.                                             */
#define DNV 4
#define DNH (DNV * (DNV + 1) / 2)
typedef struct { double v, g[DNV], h[DNH]; } dd_t;
static int dd_hidx(int nv, int i, int j) { return i * nv - (i * (i - 1)) / 2 + (j - i); }

static dd_t dd_seed(int nv, int i, double v) {
  dd_t r;
  memset(&r, 0, sizeof r);
  (void)nv;
  r.v = v;
  r.g[i] = 1.0;
  return r;
}
static dd_t dd_add(int nv, int ord, dd_t a, dd_t b) {
  dd_t r;
  r.v = a.v + b.v;
  for (int i = 0; i < nv; i++) r.g[i] = a.g[i] + b.g[i];
  if (ord > 1) for (int k = 0; k < nv * (nv + 1) / 2; k++) r.h[k] = a.h[k] + b.h[k];
  return r;
}
static dd_t dd_sub(int nv, int ord, dd_t a, dd_t b) {
  dd_t r;
  r.v = a.v - b.v;
  for (int i = 0; i < nv; i++) r.g[i] = a.g[i] - b.g[i];
  if (ord > 1) for (int k = 0; k < nv * (nv + 1) / 2; k++) r.h[k] = a.h[k] - b.h[k];
  return r;
}
static dd_t dd_mul(int nv, int ord, dd_t a, dd_t b) {
  dd_t r;
  r.v = a.v * b.v;
  for (int i = 0; i < nv; i++) r.g[i] = fma(a.g[i], b.v, a.v * b.g[i]);
  if (ord > 1)
    for (int i = 0; i < nv; i++)
      for (int j = i; j < nv; j++) {
        int k = dd_hidx(nv, i, j);
        double t = a.v * b.h[k];
        t = fma(a.g[i], b.g[j], t);
        t = fma(a.g[j], b.g[i], t);
        t = fma(a.h[k], b.v, t);
        r.h[k] = t;
      }
  return r;
}
static dd_t dd_mulc(int nv, int ord, dd_t a, double c) {
  dd_t r;
  r.v = a.v * c;
  for (int i = 0; i < nv; i++) r.g[i] = a.g[i] * c;
  if (ord > 1) for (int k = 0; k < nv * (nv + 1) / 2; k++) r.h[k] = a.h[k] * c;
  return r;
}
static dd_t dd_csub(int nv, int ord, double c, dd_t a) { /* c - a */
  dd_t r;
  r.v = c - a.v;
  for (int i = 0; i < nv; i++) r.g[i] = -a.g[i];
  if (ord > 1) for (int k = 0; k < nv * (nv + 1) / 2; k++) r.h[k] = -a.h[k];
  return r;
}
static dd_t dd_addc(dd_t a, double c) { a.v = a.v + c; return a; }

static double control_of(const problem_t *pb, const double *y) {
  switch (pb->id) {
    case PB_VDP_PMP_SMOOTH: return -y[1] * ((1.0 - y[2] * y[2]) * y[3] - y[2]) / pb->par[0];
    case PB_VDP_PMP_BANG: return (y[1] * (1.0 - y[2] * y[2]) * y[3] < 0.0) ? pb->par[2] : pb->par[1];
    case PB_VDP_AUG_BANG: return (y[1] * (1.0 - y[2] * y[2]) * y[3] < 0.0) ? pb->par[1] : pb->par[0];
    default: return 0.0;
  }
}

/* H evaluated on dual numbers; variables in y-order (p..., x...) */
static dd_t ham_mirror(const problem_t *pb, const double *y, int ord) {
  const int nv = pb->Ny;
  const double u = control_of(pb, y);
  switch (pb->id) {
    case PB_HARMONIC: {
      const double w = pb->par[0];
      dd_t p = dd_seed(nv, 0, y[0]), x = dd_seed(nv, 1, y[1]);
      dd_t t1 = dd_mul(nv, ord, p, p), t2 = dd_mul(nv, ord, x, x);
      return dd_mulc(nv, ord, dd_add(nv, ord, t1, dd_mulc(nv, ord, t2, w * w)), 0.5);
    }
    case PB_VDP: {
      const double mu = pb->par[0];
      dd_t p1 = dd_seed(nv, 0, y[0]), p2 = dd_seed(nv, 1, y[1]), x1 = dd_seed(nv, 2, y[2]), x2 = dd_seed(nv, 3, y[3]);
      dd_t a = dd_mul(nv, ord, p1, x2);
      dd_t c = dd_csub(nv, ord, 1.0, dd_mul(nv, ord, x1, x1));
      dd_t e = dd_mul(nv, ord, dd_mulc(nv, ord, c, mu), x2);
      dd_t g = dd_mul(nv, ord, p2, dd_sub(nv, ord, e, x1));
      return dd_add(nv, ord, a, g);
    }
    case PB_VDP_PMP_SMOOTH:
    case PB_VDP_PMP_BANG: {
      const double W = pb->par[0];
      dd_t p1 = dd_seed(nv, 0, y[0]), p2 = dd_seed(nv, 1, y[1]), x1 = dd_seed(nv, 2, y[2]), x2 = dd_seed(nv, 3, y[3]);
      dd_t a = dd_mul(nv, ord, p1, x2);
      dd_t pu = dd_mulc(nv, ord, p2, u);
      dd_t c = dd_csub(nv, ord, 1.0, dd_mul(nv, ord, x1, x1));
      dd_t f = dd_sub(nv, ord, dd_mul(nv, ord, c, x2), x1);
      dd_t r = dd_add(nv, ord, a, dd_mul(nv, ord, pu, f));
      return dd_addc(r, (W / 2) * u * u + 1.0);
    }
    default: { /* PB_VDP_AUG_BANG */
      dd_t p1 = dd_seed(nv, 0, y[0]), p2 = dd_seed(nv, 1, y[1]), x1 = dd_seed(nv, 2, y[2]), x2 = dd_seed(nv, 3, y[3]);
      dd_t a = dd_mul(nv, ord, p1, x2);
      dd_t c = dd_csub(nv, ord, 1.0, dd_mul(nv, ord, x1, x1));
      dd_t e = dd_mul(nv, ord, dd_mulc(nv, ord, c, u), x2);
      dd_t g = dd_mul(nv, ord, p2, dd_sub(nv, ord, e, x1));
      return dd_add(nv, ord, a, g);
    }
  }
}

static void rhs_mirror(const problem_t *pb, const double *y, double *f) {
  const int nv = pb->Ny, N = nv / 2;
  dd_t r = ham_mirror(pb, y, 1);
  for (int i = 0; i < N; i++) {
    f[i] = pb->sign * r.g[N + i];
    f[N + i] = r.g[i];
  }
}
static void jac_mirror(const problem_t *pb, const double *y, double *J) {
  const int nv = pb->Ny, N = nv / 2;
  dd_t r = ham_mirror(pb, y, 2);
  for (int rr = 0; rr < nv; rr++)
    for (int c = 0; c < nv; c++) {
      int vi = (rr < N) ? (N + rr) : (rr - N);
      int a = vi < c ? vi : c, b = vi < c ? c : vi;
      double val = r.h[dd_hidx(nv, a, b)];
      J[rr * nv + c] = (rr < N) ? pb->sign * val : val;
    }
}

static void rhs(const problem_t *pb, double t, const double *y, double *f) {
  (void)t;
  const double sg = pb->sign;
  if (g_mirror && pb->id == PB_MLP) { mlp_eval_mirror(pb, y, f, NULL); return; }
  if (g_mirror) { rhs_mirror(pb, y, f); return; }
  switch (pb->id) {
    case PB_HARMONIC: {
      double w = pb->par[0];
      f[0] = sg * (w * w * y[1]);
      f[1] = y[0];
    } break;
    case PB_VDP: {
      double mu = pb->par[0], p1 = y[0], p2 = y[1], x1 = y[2], x2 = y[3];
      f[0] = sg * (p2 * (-2.0 * mu * x1 * x2 - 1.0));
      f[1] = sg * (p1 + p2 * (mu * (1.0 - x1 * x1)));
      f[2] = x2;
      f[3] = mu * (1.0 - x1 * x1) * x2 - x1;
    } break;
    case PB_VDP_PMP_SMOOTH:
    case PB_VDP_PMP_BANG: {
      double W = pb->par[0], p1 = y[0], p2 = y[1], x1 = y[2], x2 = y[3], u;
      if (pb->id == PB_VDP_PMP_SMOOTH) u = -p2 * ((1.0 - x1 * x1) * x2 - x1) / W;
      else u = bang(p2, x1, x2, pb->par[1], pb->par[2]);
      f[0] = sg * (p2 * u * (-2.0 * x1 * x2 - 1.0));
      f[1] = sg * (p1 + p2 * u * (1.0 - x1 * x1));
      f[2] = x2;
      f[3] = u * ((1.0 - x1 * x1) * x2 - x1);
    } break;
    case PB_VDP_AUG_BANG: {
      double p1 = y[0], p2 = y[1], x1 = y[2], x2 = y[3];
      double u = bang(p2, x1, x2, pb->par[0], pb->par[1]);
      f[0] = sg * (p2 * u * (-2.0 * x1) * x2 - p2);
      f[1] = sg * (p1 + p2 * u * (1.0 - x1 * x1));
      f[2] = x2;
      f[3] = u * (1.0 - x1 * x1) * x2 - x1;
    } break;
    case PB_MLP:
      mlp_eval(pb, y, f, NULL);
      break;
  }
}

static void jacobian(const problem_t *pb, double t, const double *y, double *J /*[Ny*Ny] row-major*/) {
  (void)t;
  const int n = pb->Ny;
  const double sg = pb->sign;
  if (g_mirror && pb->id == PB_MLP) { double fm[NYMAX]; mlp_eval_mirror(pb, y, fm, J); return; }
  if (g_mirror) { jac_mirror(pb, y, J); return; }
  memset(J, 0, sizeof(double) * n * n);
  switch (pb->id) {
    case PB_HARMONIC: {
      double w = pb->par[0];
      J[0 * 2 + 1] = sg * w * w;
      J[1 * 2 + 0] = 1.0;
    } break;
    case PB_VDP: {
      double mu = pb->par[0], p2 = y[1], x1 = y[2], x2 = y[3];
      J[0 * 4 + 1] = sg * (-2.0 * mu * x1 * x2 - 1.0);
      J[0 * 4 + 2] = sg * (p2 * (-2.0 * mu * x2));
      J[0 * 4 + 3] = sg * (p2 * (-2.0 * mu * x1));
      J[1 * 4 + 0] = sg * 1.0;
      J[1 * 4 + 1] = sg * (mu * (1.0 - x1 * x1));
      J[1 * 4 + 2] = sg * (p2 * mu * (-2.0 * x1));
      J[2 * 4 + 3] = 1.0;
      J[3 * 4 + 2] = -2.0 * mu * x1 * x2 - 1.0;
      J[3 * 4 + 3] = mu * (1.0 - x1 * x1);
    } break;
    case PB_VDP_PMP_SMOOTH:
    case PB_VDP_PMP_BANG: {
      double W = pb->par[0], p2 = y[1], x1 = y[2], x2 = y[3], u;
      if (pb->id == PB_VDP_PMP_SMOOTH) u = -p2 * ((1.0 - x1 * x1) * x2 - x1) / W;
      else u = bang(p2, x1, x2, pb->par[1], pb->par[2]);
      J[0 * 4 + 1] = sg * (u * (-2.0 * x1 * x2 - 1.0));
      J[0 * 4 + 2] = sg * (p2 * u * (-2.0 * x2));
      J[0 * 4 + 3] = sg * (p2 * u * (-2.0 * x1));
      J[1 * 4 + 0] = sg * 1.0;
      J[1 * 4 + 1] = sg * (u * (1.0 - x1 * x1));
      J[1 * 4 + 2] = sg * (p2 * u * (-2.0 * x1));
      J[2 * 4 + 3] = 1.0;
      J[3 * 4 + 2] = u * (-2.0 * x1 * x2 - 1.0);
      J[3 * 4 + 3] = u * (1.0 - x1 * x1);
    } break;
    case PB_VDP_AUG_BANG: {
      double p2 = y[1], x1 = y[2], x2 = y[3];
      double u = bang(p2, x1, x2, pb->par[0], pb->par[1]);
      J[0 * 4 + 1] = sg * (u * (-2.0 * x1) * x2 - 1.0);
      J[0 * 4 + 2] = sg * (-p2 * u * 2.0 * x2);
      J[0 * 4 + 3] = sg * (-p2 * u * 2.0 * x1);
      J[1 * 4 + 0] = sg * 1.0;
      J[1 * 4 + 1] = sg * (u * (1.0 - x1 * x1));
      J[1 * 4 + 2] = sg * (p2 * u * (-2.0 * x1));
      J[2 * 4 + 3] = 1.0;
      J[3 * 4 + 2] = u * (-2.0 * x1 * x2) - 1.0;
      J[3 * 4 + 3] = u * (1.0 - x1 * x1);
    } break;
    case PB_MLP: {
      double f[NYMAX];
      mlp_eval(pb, y, f, J);
    } break;
  }
}

/* MLP Hamiltonian H(v), v=[x;p] (2N inputs), 3 tanh hidden layers of `width`, scalar output.
 * gradient g = dH/dv by back-propagation, Hessian Hvv by forward-over-reverse (one tangent per
 * input).  ydot = [sg*dH/dx ; dH/dp], J = d ydot/d y with y=[p;x].                         */
static void mlp_eval(const problem_t *pb, const double *y, double *f, double *J) {
  const int N = (int)pb->par[0], Wd = (int)pb->par[1], n = 2 * N;
  const double *W1 = pb->par + 2, *b1 = W1 + (size_t)Wd * n;
  const double *W2 = b1 + Wd, *b2 = W2 + (size_t)Wd * Wd;
  const double *W3 = b2 + Wd, *b3 = W3 + (size_t)Wd * Wd;
  const double *W4 = b3 + Wd;
  double v[NYMAX];
  for (int i = 0; i < N; i++) { v[i] = y[N + i]; v[N + i] = y[i]; }   /* v=[x;p] */
  double *a1 = (double *)malloc(sizeof(double) * Wd * 12);
  double *a2 = a1 + Wd, *a3 = a2 + Wd, *d1 = a3 + Wd, *d2 = d1 + Wd, *d3 = d2 + Wd;
  double *ta1 = d3 + Wd, *ta2 = ta1 + Wd, *ta3 = ta2 + Wd, *td1 = ta3 + Wd, *td2 = td1 + Wd, *td3 = td2 + Wd;
  for (int i = 0; i < Wd; i++) { double s = b1[i]; for (int j = 0; j < n; j++) s += W1[(size_t)i * n + j] * v[j]; a1[i] = tanh(s); }
  for (int i = 0; i < Wd; i++) { double s = b2[i]; for (int j = 0; j < Wd; j++) s += W2[(size_t)i * Wd + j] * a1[j]; a2[i] = tanh(s); }
  for (int i = 0; i < Wd; i++) { double s = b3[i]; for (int j = 0; j < Wd; j++) s += W3[(size_t)i * Wd + j] * a2[j]; a3[i] = tanh(s); }
  /* reverse: d3 = W4 .* (1-a3^2); d2 = (W3^T d3) .* (1-a2^2); d1 = (W2^T d2) .* (1-a1^2); g = W1^T d1 */
  for (int i = 0; i < Wd; i++) d3[i] = W4[i] * (1.0 - a3[i] * a3[i]);
  for (int j = 0; j < Wd; j++) { double s = 0; for (int i = 0; i < Wd; i++) s += W3[(size_t)i * Wd + j] * d3[i]; d2[j] = s * (1.0 - a2[j] * a2[j]); }
  for (int j = 0; j < Wd; j++) { double s = 0; for (int i = 0; i < Wd; i++) s += W2[(size_t)i * Wd + j] * d2[i]; d1[j] = s * (1.0 - a1[j] * a1[j]); }
  double g[NYMAX];
  for (int j = 0; j < n; j++) { double s = 0; for (int i = 0; i < Wd; i++) s += W1[(size_t)i * n + j] * d1[i]; g[j] = s; }
  /* g = [dH/dx ; dH/dp];  ydot = [sg dH/dx ; dH/dp] */
  for (int i = 0; i < N; i++) { f[i] = pb->sign * g[i]; f[N + i] = g[N + i]; }
  if (J) {
    double Hvv[NYMAX * NYMAX];
    for (int k = 0; k < n; k++) {   /* tangent direction e_k in v */
      for (int i = 0; i < Wd; i++) ta1[i] = (1.0 - a1[i] * a1[i]) * W1[(size_t)i * n + k];
      for (int i = 0; i < Wd; i++) { double s = 0; for (int j = 0; j < Wd; j++) s += W2[(size_t)i * Wd + j] * ta1[j]; ta2[i] = (1.0 - a2[i] * a2[i]) * s; }
      for (int i = 0; i < Wd; i++) { double s = 0; for (int j = 0; j < Wd; j++) s += W3[(size_t)i * Wd + j] * ta2[j]; ta3[i] = (1.0 - a3[i] * a3[i]) * s; }
      for (int i = 0; i < Wd; i++) td3[i] = W4[i] * (-2.0 * a3[i] * ta3[i]);
      for (int j = 0; j < Wd; j++) {
        double s = 0, ts = 0;
        for (int i = 0; i < Wd; i++) { s += W3[(size_t)i * Wd + j] * d3[i]; ts += W3[(size_t)i * Wd + j] * td3[i]; }
        td2[j] = ts * (1.0 - a2[j] * a2[j]) + s * (-2.0 * a2[j] * ta2[j]);
      }
      for (int j = 0; j < Wd; j++) {
        double s = 0, ts = 0;
        for (int i = 0; i < Wd; i++) { s += W2[(size_t)i * Wd + j] * d2[i]; ts += W2[(size_t)i * Wd + j] * td2[i]; }
        td1[j] = ts * (1.0 - a1[j] * a1[j]) + s * (-2.0 * a1[j] * ta1[j]);
      }
      for (int j = 0; j < n; j++) { double s = 0; for (int i = 0; i < Wd; i++) s += W1[(size_t)i * n + j] * td1[i]; Hvv[j * n + k] = s; }
    }
    /* map v-index -> y-index: v_i (x_i) = y_{N+i}; v_{N+i} (p_i) = y_i.
     * ydot_r: r<N  -> sg*g[r]   (g index r = x_r);  r>=N -> g[r] (p_{r-N})
     * column c of y: c<N -> p_c = v_{N+c}; c>=N -> x_{c-N} = v_{c-N}                     */
    for (int r = 0; r < n; r++)
      for (int c = 0; c < n; c++) {
        int vc = (c < N) ? (N + c) : (c - N);
        double val = Hvv[r * n + vc];
        J[r * n + c] = (r < N) ? pb->sign * val : val;
      }
  }
  free(a1);
}


/* Mirror of the GPU's MLP kernels (csrc/mlp_kernels.cuh), operation by operation: every dot
 * product is ONE fma chain in increasing index order (forward: seeded with the bias; transposed
 * products: seeded with 0.0), tanh = det_tanh, 1 - a^2 = fma(-a, a, 1).                     */
static void mlp_eval_mirror(const problem_t *pb, const double *y, double *f, double *J) {
  const int N = (int)pb->par[0], Wd = (int)pb->par[1], n = 2 * N;
  const double *W1 = pb->par + 2, *b1 = W1 + (size_t)Wd * n;
  const double *W2 = b1 + Wd, *b2 = W2 + (size_t)Wd * Wd;
  const double *W3 = b2 + Wd, *b3 = W3 + (size_t)Wd * Wd;
  const double *W4 = b3 + Wd;
  double v[NYMAX];
  for (int i = 0; i < N; i++) { v[i] = y[N + i]; v[N + i] = y[i]; }   /* v=[x;p] */
  double *a1 = (double *)malloc(sizeof(double) * Wd * 17);
  double *a2 = a1 + Wd, *a3 = a2 + Wd, *d1 = a3 + Wd, *d2 = d1 + Wd, *d3 = d2 + Wd;
  double *ta1 = d3 + Wd, *ta2 = ta1 + Wd, *ta3 = ta2 + Wd, *td1 = ta3 + Wd, *td2 = td1 + Wd, *td3 = td2 + Wd;
  double *q1 = td3 + Wd, *q2 = q1 + Wd, *e1 = q2 + Wd, *e2 = e1 + Wd, *e3 = e2 + Wd;
  for (int i = 0; i < Wd; i++) { double s = b1[i]; for (int j = 0; j < n; j++) s = fma(W1[(size_t)i * n + j], v[j], s); a1[i] = det_tanh(s); }
  for (int i = 0; i < Wd; i++) { double s = b2[i]; for (int j = 0; j < Wd; j++) s = fma(W2[(size_t)i * Wd + j], a1[j], s); a2[i] = det_tanh(s); }
  for (int i = 0; i < Wd; i++) { double s = b3[i]; for (int j = 0; j < Wd; j++) s = fma(W3[(size_t)i * Wd + j], a2[j], s); a3[i] = det_tanh(s); }
  for (int i = 0; i < Wd; i++) { e1[i] = fma(-a1[i], a1[i], 1.0); e2[i] = fma(-a2[i], a2[i], 1.0); e3[i] = fma(-a3[i], a3[i], 1.0); }
  for (int i = 0; i < Wd; i++) d3[i] = W4[i] * e3[i];
  for (int j = 0; j < Wd; j++) { double s = 0.0; for (int i = 0; i < Wd; i++) s = fma(W3[(size_t)i * Wd + j], d3[i], s); q2[j] = s; d2[j] = s * e2[j]; }
  for (int j = 0; j < Wd; j++) { double s = 0.0; for (int i = 0; i < Wd; i++) s = fma(W2[(size_t)i * Wd + j], d2[i], s); q1[j] = s; d1[j] = s * e1[j]; }
  double g[NYMAX];
  for (int j = 0; j < n; j++) { double s = 0.0; for (int i = 0; i < Wd; i++) s = fma(W1[(size_t)i * n + j], d1[i], s); g[j] = s; }
  for (int i = 0; i < N; i++) { f[i] = pb->sign * g[i]; f[N + i] = g[N + i]; }
  if (J) {
    double Hvv[NYMAX * NYMAX];
    for (int k = 0; k < n; k++) {   /* tangent direction e_k in v */
      for (int i = 0; i < Wd; i++) ta1[i] = e1[i] * W1[(size_t)i * n + k];
      for (int i = 0; i < Wd; i++) { double s = 0.0; for (int j = 0; j < Wd; j++) s = fma(W2[(size_t)i * Wd + j], ta1[j], s); ta2[i] = e2[i] * s; }
      for (int i = 0; i < Wd; i++) { double s = 0.0; for (int j = 0; j < Wd; j++) s = fma(W3[(size_t)i * Wd + j], ta2[j], s); ta3[i] = e3[i] * s; }
      for (int i = 0; i < Wd; i++) td3[i] = W4[i] * ((-2.0 * a3[i]) * ta3[i]);
      for (int j = 0; j < Wd; j++) {
        double ts = 0.0;
        for (int i = 0; i < Wd; i++) ts = fma(W3[(size_t)i * Wd + j], td3[i], ts);
        td2[j] = fma(q2[j], (-2.0 * a2[j]) * ta2[j], ts * e2[j]);
      }
      for (int j = 0; j < Wd; j++) {
        double ts = 0.0;
        for (int i = 0; i < Wd; i++) ts = fma(W2[(size_t)i * Wd + j], td2[i], ts);
        td1[j] = fma(q1[j], (-2.0 * a1[j]) * ta1[j], ts * e1[j]);
      }
      for (int j = 0; j < n; j++) { double s = 0.0; for (int i = 0; i < Wd; i++) s = fma(W1[(size_t)i * n + j], td1[i], s); Hvv[j * n + k] = s; }
    }
    for (int r = 0; r < n; r++)
      for (int c = 0; c < n; c++) {
        int vc = (c < N) ? (N + c) : (c - N);
        double val = Hvv[r * n + vc];
        J[r * n + c] = (r < N) ? pb->sign * val : val;
      }
  }
  free(a1);
}

/* ------------------------------------------------------------------ LU (zgetf2 rules) */
typedef struct {
  double a[NYMAX * NYMAX];   /* L\U, row-major */
  double rd[NYMAX];          /* reciprocal of U's diagonal */
  int piv[NYMAX];
} rlu_t;

typedef struct {
  double re[NYMAX * NYMAX], im[NYMAX * NYMAX];
  double rdre[NYMAX], rdim[NYMAX];
  int piv[NYMAX];
} zlu_t;

static void rlu_factor(int n, rlu_t *L) {
  double *a = L->a;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(a[k * n + k]);
    for (int i = k + 1; i < n; i++) {
      double v = fabs(a[i * n + k]);
      if (v > best) { best = v; p = i; }
    }
    L->piv[k] = p;
    if (p != k)
      for (int j = 0; j < n; j++) { double tmp = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = tmp; }
    double r = 1.0 / a[k * n + k];
    L->rd[k] = r;
    for (int i = k + 1; i < n; i++) {
      double l = a[i * n + k] * r;
      a[i * n + k] = l;
      for (int j = k + 1; j < n; j++) a[i * n + j] = FMA(-l, a[k * n + j], a[i * n + j]);
    }
  }
}

static void rlu_solve(int n, const rlu_t *L, double *b) {
  const double *a = L->a;
  for (int k = 0; k < n; k++) {
    int p = L->piv[k];
    if (p != k) { double tmp = b[k]; b[k] = b[p]; b[p] = tmp; }
  }
  for (int i = 1; i < n; i++) {
    double s = b[i];
    for (int j = 0; j < i; j++) s = FMA(-a[i * n + j], b[j], s);
    b[i] = s;
  }
  if (g_mirror && n == NYMAX) {
    /* Ny = 16 (MLP Hamiltonian path, csrc/lu16.cuh): column-oriented back substitution, the order a
     * lane-parallel triangular solve (and reference-BLAS dtrsm) produces: unknown j is finished, then
     * every row above it is updated                                                              */
    for (int j = n - 1; j >= 0; j--) {
      const double xj = b[j] * L->rd[j];
      b[j] = xj;
      for (int i = 0; i < j; i++) b[i] = fma(-a[i * n + j], xj, b[i]);
    }
    return;
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int j = i + 1; j < n; j++) s = FMA(-a[i * n + j], b[j], s);
    b[i] = s * L->rd[i];
  }
}

static void zlu_factor(int n, zlu_t *L) {
  double *ar = L->re, *ai = L->im;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(ar[k * n + k]) + fabs(ai[k * n + k]);
    for (int i = k + 1; i < n; i++) {
      double v = fabs(ar[i * n + k]) + fabs(ai[i * n + k]);
      if (v > best) { best = v; p = i; }
    }
    L->piv[k] = p;
    if (p != k)
      for (int j = 0; j < n; j++) {
        double tmp = ar[k * n + j]; ar[k * n + j] = ar[p * n + j]; ar[p * n + j] = tmp;
        tmp = ai[k * n + j]; ai[k * n + j] = ai[p * n + j]; ai[p * n + j] = tmp;
      }
    double c = ar[k * n + k], d = ai[k * n + k];
    double inv = 1.0 / FMA(d, d, c * c);
    double rr = c * inv, ri = -d * inv;
    L->rdre[k] = rr; L->rdim[k] = ri;
    for (int i = k + 1; i < n; i++) {
      double xr = ar[i * n + k], xi = ai[i * n + k];
      double lr = FMA(-xi, ri, xr * rr), li = FMA(xi, rr, xr * ri);
      ar[i * n + k] = lr; ai[i * n + k] = li;
      for (int j = k + 1; j < n; j++) {
        double ur = ar[k * n + j], ui = ai[k * n + j];
        if (g_mirror) {
          ar[i * n + j] = fma(li, ui, fma(-lr, ur, ar[i * n + j]));
          ai[i * n + j] = fma(-li, ur, fma(-lr, ui, ai[i * n + j]));
        } else {
          ar[i * n + j] = ar[i * n + j] - (lr * ur - li * ui);
          ai[i * n + j] = ai[i * n + j] - (lr * ui + li * ur);
        }
      }
    }
  }
}

static void zlu_solve(int n, const zlu_t *L, double *br, double *bi) {
  const double *ar = L->re, *ai = L->im;
  for (int k = 0; k < n; k++) {
    int p = L->piv[k];
    if (p != k) {
      double tmp = br[k]; br[k] = br[p]; br[p] = tmp;
      tmp = bi[k]; bi[k] = bi[p]; bi[p] = tmp;
    }
  }
  for (int i = 1; i < n; i++) {
    double sr = br[i], si = bi[i];
    for (int j = 0; j < i; j++) {
      double lr = ar[i * n + j], li = ai[i * n + j];
      if (g_mirror) {
        sr = fma(li, bi[j], fma(-lr, br[j], sr));
        si = fma(-li, br[j], fma(-lr, bi[j], si));
      } else {
        sr = sr - (lr * br[j] - li * bi[j]);
        si = si - (lr * bi[j] + li * br[j]);
      }
    }
    br[i] = sr; bi[i] = si;
  }
  if (g_mirror && n == NYMAX) {   /* column-oriented, see rlu_solve */
    for (int j = n - 1; j >= 0; j--) {
      const double rr = L->rdre[j], ri = L->rdim[j];
      const double xr = fma(-bi[j], ri, br[j] * rr), xi = fma(bi[j], rr, br[j] * ri);
      br[j] = xr; bi[j] = xi;
      for (int i = 0; i < j; i++) {
        const double ur = ar[i * n + j], ui = ai[i * n + j];
        const double nr = fma(ui, xi, fma(-ur, xr, br[i]));
        bi[i] = fma(-ui, xr, fma(-ur, xi, bi[i]));
        br[i] = nr;
      }
    }
    return;
  }
  for (int i = n - 1; i >= 0; i--) {
    double sr = br[i], si = bi[i];
    for (int j = i + 1; j < n; j++) {
      double ur = ar[i * n + j], ui = ai[i * n + j];
      if (g_mirror) {
        sr = fma(ui, bi[j], fma(-ur, br[j], sr));
        si = fma(-ui, br[j], fma(-ur, bi[j], si));
      } else {
        sr = sr - (ur * br[j] - ui * bi[j]);
        si = si - (ur * bi[j] + ui * br[j]);
      }
    }
    double rr = L->rdre[i], ri = L->rdim[i];
    br[i] = FMA(-si, ri, sr * rr);
    bi[i] = FMA(si, rr, sr * ri);
  }
}

/* ------------------------------------------------------------------ tables */
typedef struct {
  int s, Nit;
  const double *C, *Dd, *ValP;
  double T[SMAX][SMAX], TI[SMAX][SMAX];
} tab_t;

static void set_tables(tab_t *tb, int s, int MaxNbrNewton) {
  tb->s = s;
  const double *T, *TI;
  switch (s) {
    case 1:
      tb->C = RADAU_C1; tb->Dd = RADAU_Dd1; tb->ValP = RADAU_ValP1; tb->Nit = MaxNbrNewton - 3;
      T = &RADAU_T1[0][0]; TI = &RADAU_TI1[0][0];
      break;
    case 3:
      tb->C = RADAU_C3; tb->Dd = RADAU_Dd3; tb->ValP = RADAU_ValP3; tb->Nit = MaxNbrNewton;
      T = &RADAU_T3[0][0]; TI = &RADAU_TI3[0][0];
      break;
    case 5:
      tb->C = RADAU_C5; tb->Dd = RADAU_Dd5; tb->ValP = RADAU_ValP5; tb->Nit = MaxNbrNewton + 5;
      T = &RADAU_T5[0][0]; TI = &RADAU_TI5[0][0];
      break;
    default:
      tb->C = RADAU_C7; tb->Dd = RADAU_Dd7; tb->ValP = RADAU_ValP7; tb->Nit = MaxNbrNewton + 10;
      T = &RADAU_T7[0][0]; TI = &RADAU_TI7[0][0];
      break;
  }
  for (int i = 0; i < s; i++)
    for (int j = 0; j < s; j++) {
      tb->T[i][j] = T[i * s + j];
      tb->TI[i][j] = TI[i * s + j];
    }
}

static int any_nan(const double *v, int n) {
  for (int i = 0; i < n; i++) if (isnan(v[i])) return 1;
  return 0;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* stats index: 0 FcnNbr 1 JacNbr 2 DecompNbr 3 SolveNbr 4 StepNbr 5 AccptNbr 6 StepRejNbr 7 NewtRejNbr */
enum { ST_FCN, ST_JAC, ST_DEC, ST_SOL, ST_STEP, ST_ACC, ST_REJ, ST_NREJ };

typedef struct {
  rlu_t L0;
  zlu_t Lc[(SMAX - 1) / 2];
} lus_t;

static void decomrc(int n, const tab_t *tb, double h, const double *Jac, lus_t *lu) {
  double rh = 1.0 / h;
  double valp[SMAX];
  for (int q = 0; q < tb->s; q++) valp[q] = tb->ValP[q] * rh;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) lu->L0.a[i * n + j] = -Jac[i * n + j] + ((i == j) ? valp[0] : 0.0);
  rlu_factor(n, &lu->L0);
  for (int q = 1; q <= (tb->s - 1) / 2; q++) {
    zlu_t *Z = &lu->Lc[q - 1];
    double al = valp[2 * q - 1], be = valp[2 * q];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        Z->re[i * n + j] = ((i == j) ? al : 0.0) - Jac[i * n + j];
        Z->im[i * n + j] = (i == j) ? be : 0.0;
      }
    zlu_factor(n, Z);
  }
}

/* z, w : [s][n] (stage-major) */
static void solvrad(int n, const tab_t *tb, double h, const lus_t *lu, double z[SMAX][NYMAX], double w[SMAX][NYMAX]) {
  double rh = 1.0 / h;
  double valp[SMAX];
  for (int q = 0; q < tb->s; q++) valp[q] = tb->ValP[q] * rh;
  for (int i = 0; i < n; i++) z[0][i] = FMA(-valp[0], w[0][i], z[0][i]);
  rlu_solve(n, &lu->L0, z[0]);
  for (int q = 1; q <= (tb->s - 1) / 2; q++) {
    int q2 = 2 * q - 1, q3 = 2 * q;
    double z2[NYMAX], z3[NYMAX];
    for (int i = 0; i < n; i++) {
      z2[i] = FMA(valp[q3], w[q3][i], FMA(-valp[q2], w[q2][i], z[q2][i]));
      z3[i] = FMA(-valp[q2], w[q3][i], FMA(-valp[q3], w[q2][i], z[q3][i]));
    }
    zlu_solve(n, &lu->Lc[q - 1], z2, z3);
    for (int i = 0; i < n; i++) { z[q2][i] = z2[i]; z[q3][i] = z3[i]; }
  }
}

static double scaled_norm(int n, const double *v, const double *Scal) {
  double s = 0.0;
  for (int i = 0; i < n; i++) { double q = g_mirror ? v[i] * (1.0 / Scal[i]) : v[i] / Scal[i]; s = FMA(q, q, s); }
  return sqrt(s);
}

static double estrad(const problem_t *pb, const tab_t *tb, double h, const lus_t *lu, double z[SMAX][NYMAX],
                     const double *Scal, const double *f0, int First, int Reject, double t, const double *y,
                     long long *st) {
  int n = pb->Ny;
  double rh = 1.0 / h, temp[NYMAX], ev[NYMAX];
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int q = 0; q < tb->s; q++) s = FMA(z[q][i], tb->Dd[q] * rh, s);
    temp[i] = s;
    ev[i] = f0[i] + s;
  }
  rlu_solve(n, &lu->L0, ev);
  double err = scaled_norm(n, ev, Scal);
  err = g_mirror ? fmax(err * (1.0 / sqrt((double)n)), 1e-10) : fmax(err / sqrt((double)n), 1e-10);
  if (err >= 1.0 && (First || Reject)) {
    double ya[NYMAX];
    for (int i = 0; i < n; i++) ya[i] = y[i] + err;
    rhs(pb, t, ya, ev);
    st[ST_FCN] += 1;
    for (int i = 0; i < n; i++) ev[i] = ev[i] + temp[i];
    rlu_solve(n, &lu->L0, ev);
    err = scaled_norm(n, ev, Scal);
    err = g_mirror ? fmax(err * (1.0 / sqrt((double)n)), 1e-10) : fmax(err / sqrt((double)n), 1e-10);
  }
  return err;
}

/* per-pass log of oracle_radau_trace2 (single-threaded diagnostic): rows of 8 doubles =
 * [StepNbr, t, h of the attempt, NbrStg, Newt, Theta, err (-1: Newton failed), FacConv] */
/* reference-variant switches, see oracle_set_quirks */
static int g_quirks = 1;
#define Q8_GUS_UNGATED 1   /* RadauTe applies the Gustafsson limiter on every accepted step with AccptNbr > 1, also when the
                              order just changed (radaute_impl.hpp:1092-1100: m1_12_1_1_1 lacks `& ~ChangeFlag`); hacc / erracc
                              are still only updated under `Gustafsson & ~ChangeFlag`.  MATLAB radau.m:1066-1076 gates both. */
static double *g_plog = 0; static int g_plog_cap = 0, g_plog_n = 0;
/* one trajectory; returns res code */
/* dense output rows of one trajectory (radau.m:1150-1168 OutFlag 3); tsp == NULL: final state only */
typedef struct {
  const double *tsp; int ntspan; double *tout, *yout; int nout;
  /* mode 3: tspan points (default, mode == 0 means 3); 1 / 2: every accepted step (+ refine - 1 interpolated points),
   * radaute_impl.hpp:1166-1183 / radau.m:1122-1143, stored while nout < cap */
  int mode, cap, refine;
  /* events E_k = A[k].y - b[k] (radau.m:1110-1121, EventZeroFcn :1494-1618) */
  int ev_K; const double *ev_A, *ev_b; const int *ev_stop, *ev_slope;
  double *te, *ye; int *ie; int nev, ecap;
} dense_t;

static double eps_of(double t) {
  double a = fabs(t);
  if (a < 2.2250738585072014e-308) return 4.9406564584124654e-324;
  int e; frexp(a, &e);            /* a = f 2^e, f in [0.5,1)  ->  spacing 2^(e-53) */
  return ldexp(1.0, e - 53);
}

static void ntrp(int n, int s, const tab_t *tb, double cont[SMAX][NYMAX], double t, const double *y, double h_old,
                 double tn, double *out) {
  const double sfr = (tn - t) / h_old;
  for (int i = 0; i < n; i++) {
    double yi = (sfr - (tb->C[0] - 1.0)) * cont[s - 1][i];
    for (int q = 1; q < s; q++) yi = (sfr - (tb->C[q] - 1.0)) * (yi + cont[s - 1 - q][i]);
    out[i] = yi + y[i];
  }
}

/* EventZeroFcn (radau.m:1494-1618) after an accepted step from (t1, E1v) to (t, y); returns 1 on a terminal event */
static int locate_events(int n, int s, const tab_t *tb, double cont[SMAX][NYMAX], dense_t *dn, const double *E1v, double t1,
                         double *t_io, double *y, double h_old) {
  const double t2 = *t_io;
  double tol = 65536.0 * fmax(EPS, fmax(eps_of(t2), eps_of(t1)));
  tol = fmin(tol, fabs(t2 - t1));
  int found = 0, stop = 0, ibest = 0;
  double tbest = 0.0, ybest[NYMAX];
  for (int k = 0; k < dn->ev_K; k++) {
    double E2 = -dn->ev_b[k];
    for (int i = 0; i < n; i++) E2 = FMA(dn->ev_A[k * n + i], y[i], E2);
    double t1N = t1, t2N = t2, E1 = E1v[k];
    const double p12 = (E2 - E1) / (t2N - t1N);
    if (!(E1 * E2 < 0 && p12 * dn->ev_slope[k] >= 0)) continue;
    double tNew = t2N, ENew = E2, yNew[NYMAX];
    for (int i = 0; i < n; i++) yNew[i] = y[i];
    for (int iter = 1; iter < 50; iter++) {
      const int tRel = fabs(t1N - t2N) * tol < fmax(fabs(t1N), fabs(t2N));
      const int tAbs = fabs(t1N - t2N) < tol;
      if ((fabs(ENew) < tol && tRel && tAbs) || fabs(ENew) == 0.0) break;
      if (fabs(E1) < 200.0 * tol || fabs(E2) < 200.0 * tol) tNew = 0.5 * (t1N + t2N);
      else tNew = tNew + (-ENew / (E2 - E1) * (t2N - t1N));
      ntrp(n, s, tb, cont, t2, y, h_old, tNew, yNew);
      ENew = -dn->ev_b[k];
      for (int i = 0; i < n; i++) ENew = FMA(dn->ev_A[k * n + i], yNew[i], ENew);
      if (ENew * E2 < 0) { t1N = t2N; E1 = E2; }
      else E1 = E1 * E2 / (E2 + ENew);
      t2N = tNew;
      E2 = ENew;
    }
    if (!found || (t1 < t2 ? tNew < tbest : tNew > tbest)) {
      found = 1; tbest = tNew; ibest = k + 1; stop = dn->ev_stop[k] != 0;
      for (int i = 0; i < n; i++) ybest[i] = yNew[i];
    }
  }
  if (!found) return 0;
  if (dn->nev < dn->ecap) {
    dn->te[dn->nev] = tbest; dn->ie[dn->nev] = ibest;
    for (int i = 0; i < n; i++) dn->ye[dn->nev * n + i] = ybest[i];
  }
  dn->nev += 1;
  if (stop) {
    *t_io = tbest;
    for (int i = 0; i < n; i++) y[i] = ybest[i];
  }
  return stop;
}

static int radau_one(const problem_t *pb, const oracle_opts *op, double t0, double tfinal, const double *y0,
                     double *y_out, double *t_out, double *h_out, long long *st, int *nstg_out,
                     double *rec_h, int rec_cap, int *rec_n, dense_t *dn) {
  const int n = pb->Ny;
  double y[NYMAX], f0[NYMAX], Scal[NYMAX], RelTol1[NYMAX], AbsTol1[NYMAX], QuotTol[NYMAX];
  double Jac[NYMAX * NYMAX];
  double z[SMAX][NYMAX], w[SMAX][NYMAX], cont[SMAX][NYMAX], fq[SMAX][NYMAX];
  lus_t lu;
  tab_t tb;
  int res = 0;
  memset(st, 0, 8 * sizeof(long long));
  memset(z, 0, sizeof z); memset(w, 0, sizeof w); memset(cont, 0, sizeof cont);
  for (int i = 0; i < n; i++) y[i] = y0[i];
  double t = t0;
  double PosNeg = (tfinal - t0) >= 0 ? 1.0 : -1.0;
  const double Thet = op->JacRecompute, Safe = op->Safe, Quot1 = op->Quot1, Quot2 = op->Quot2;
  const double FacL = 1.0 / op->FacL, FacR = 1.0 / op->FacR;
  int NbrStg = op->NbrStg;
  const int MinNbrStg = op->MinNbrStg, MaxNbrStg = op->MaxNbrStg;
  const int Variab = (MaxNbrStg - MinNbrStg) != 0;
  set_tables(&tb, NbrStg, op->MaxNbrNewton);
  int UnExpStepRej = 0, UnExpNewtRej = 0, Keep = 0, ChangeFlag = 0, ChangeNbr = 0;
  double Theta = 0.0, Thetat = 0.0;
  double hmaxn = fabs(tfinal - t0);
  double h = op->InitialStep;
  if (fabs(h) <= 10 * EPS) h = 1e-6;
  h = PosNeg * fmin(fabs(h), hmaxn);
  double h_old = h, hopt = h;
  int Last;
  if ((t + h * 1.0001 - tfinal) * PosNeg >= 0) { h = tfinal - t; Last = 1; } else Last = 0;
  double FacConv = 1.0;
  int Reject = 0, First = 1;
  double ExpmNs = (NbrStg + 1.0) / (2.0 * NbrStg);
  for (int i = 0; i < n; i++) {
    QuotTol[i] = op->AbsTol[i] / op->RelTol[i];
    RelTol1[i] = 0.1 * pow(op->RelTol[i], ExpmNs);
    AbsTol1[i] = RelTol1[i] * QuotTol[i];
    Scal[i] = FMA(RelTol1[i], fabs(y[i]), AbsTol1[i]);
  }
  rhs(pb, t, y, f0);
  st[ST_FCN] += 1;
  if (any_nan(f0, n)) { res = 2; goto done; }
  int NeedNewJac = 1, NeedNewLU = 1, Newt = 0, finished = 0;
  double SqrtStgNy = sqrt((double)(NbrStg * n));
  double hacc = 0.0, erracc = 0.0, thqold = 0.0, OldNrm = 0.0;
  if (rec_n) *rec_n = 0;

  while (st[ST_STEP] <= op->MaxNbrStep && PosNeg * t < PosNeg * tfinal) {
    st[ST_STEP] += 1;
    FacConv = POW(fmax(FacConv, EPS), 0.8);
    if (NeedNewJac) {
      jacobian(pb, t, y, Jac);
      if (any_nan(Jac, n * n)) { res = 2; break; }
      st[ST_JAC] += 1;
      NeedNewJac = 0;
      NeedNewLU = 1;
    }
    if (Variab && !Keep) {
      ChangeNbr += 1;
      int NbrStgNew = NbrStg;
      double hquot = h / h_old;
      Thetat = fmin(10.0, fmax(Theta, Thetat * 0.5));
      if (Newt > 1 && Thetat <= op->Vitu && hquot < op->hhou && hquot > op->hhod) NbrStgNew = imin(MaxNbrStg, NbrStg + 2);
      if (Thetat >= op->Vitd || UnExpStepRej) NbrStgNew = imax(MinNbrStg, NbrStg - 2);
      if (ChangeNbr >= 1 && UnExpNewtRej) NbrStgNew = imax(MinNbrStg, NbrStg - 2);
      if (ChangeNbr <= 10) NbrStgNew = imin(NbrStg, NbrStgNew);
      ChangeFlag = NbrStg != NbrStgNew;
      UnExpNewtRej = 0;
      UnExpStepRej = 0;
      if (ChangeFlag) {
        NbrStg = NbrStgNew;
        ChangeNbr = 1;
        ExpmNs = (NbrStg + 1.0) / (2.0 * NbrStg);
        for (int i = 0; i < n; i++) {
          RelTol1[i] = 0.1 * pow(op->RelTol[i], ExpmNs);
          AbsTol1[i] = RelTol1[i] * QuotTol[i];
          Scal[i] = FMA(RelTol1[i], fabs(y[i]), AbsTol1[i]);
        }
        NeedNewLU = 1;
        SqrtStgNy = sqrt((double)(NbrStg * n));
        set_tables(&tb, NbrStg, op->MaxNbrNewton);
      }
    }
    if (NeedNewLU) {
      decomrc(n, &tb, h, Jac, &lu);
      st[ST_DEC] += 1;
      NeedNewLU = 0;
    }
    if (Variab && Keep) {
      Keep = 0;
      ChangeNbr += 1;
      if (ChangeNbr >= 10 && NbrStg < MaxNbrStg) NeedNewJac = 0;
    }
    if (0.1 * fabs(h) <= fabs(t) * EPS) { res = 1; break; }
    ExpmNs = (NbrStg + 1.0) / (2.0 * NbrStg);
    double mrt = INFINITY;
    for (int i = 0; i < n; i++) {
      QuotTol[i] = op->AbsTol[i] / op->RelTol[i];
      RelTol1[i] = 0.1 * pow(op->RelTol[i], ExpmNs);
      AbsTol1[i] = RelTol1[i] * QuotTol[i];
      Scal[i] = FMA(RelTol1[i], fabs(y[i]), AbsTol1[i]);
      mrt = fmin(mrt, RelTol1[i]);
    }
    const int s = NbrStg;
    if (First || op->Start_Newt || ChangeFlag) {
      memset(z, 0, sizeof z); memset(w, 0, sizeof w); memset(cont, 0, sizeof cont);
    } else {
      double hquot = h / h_old;
      for (int q = 0; q < s; q++) {
        double cq = tb.C[q] * hquot;
        for (int i = 0; i < n; i++) {
          double zq = (cq - tb.C[0] + 1.0) * cont[s - 1][i];
          for (int q1 = 2; q1 <= s; q1++) zq = (zq + cont[s - q1][i]) * (cq - tb.C[q1 - 1] + 1.0);
          z[q][i] = zq;
        }
      }
      for (int j = 0; j < s; j++)
        for (int i = 0; i < n; i++) {
          double acc = tb.TI[j][0] * z[0][i];
          for (int q = 1; q < s; q++) acc = FMA(tb.TI[j][q], z[q][i], acc);
          w[j][i] = acc;
        }
    }
    double FNewt = fmax(10 * EPS / mrt, fmin(0.03, pow(mrt, 1.0 / ExpmNs - 1.0)));
    if (s == 1) FNewt = fmax(10 * EPS / mrt, 0.03);
    FacConv = POW(fmax(FacConv, EPS), 0.8);
    Theta = fabs(Thet);
    Newt = 0;
    const double h_attempt = h;
    int newton_failed = 0, nan_fail = 0;
    double NewNrm = 0.0;
    for (;;) {
      Reject = 0;
      Newt += 1;
      if (Newt > tb.Nit) {
        UnExpStepRej = 1;
        st[ST_REJ] += 1; st[ST_NREJ] += 1;
        h = h * 0.5;
        Reject = 1; Last = 0; NeedNewLU = 1; newton_failed = 1;
        break;
      }
      for (int q = 0; q < s; q++) {
        double yq[NYMAX];
        for (int i = 0; i < n; i++) yq[i] = y[i] + z[q][i];
        rhs(pb, t + tb.C[q] * h, yq, fq[q]);
        if (any_nan(fq[q], n)) nan_fail = 1;
      }
      if (nan_fail) break;
      st[ST_FCN] += s;
      for (int j = 0; j < s; j++)
        for (int i = 0; i < n; i++) {
          double acc = tb.TI[j][0] * fq[0][i];
          for (int q = 1; q < s; q++) acc = FMA(tb.TI[j][q], fq[q][i], acc);
          z[j][i] = acc;
        }
      solvrad(n, &tb, h, &lu, z, w);
      st[ST_SOL] += 1;
      NewNrm = 0.0;
      for (int q = 0; q < s; q++) NewNrm += scaled_norm(n, z[q], Scal);
      NewNrm = g_mirror ? NewNrm * (1.0 / SqrtStgNy) : NewNrm / SqrtStgNy;
      if (!isfinite(NewNrm)) { nan_fail = 1; break; }
      if (Newt > 1 && Newt < tb.Nit) {
        double thq = NewNrm / OldNrm;
        if (Newt == 2) Theta = thq; else Theta = sqrt(thq * thqold);
        thqold = thq;
        if (Theta < 0.99) {
          FacConv = Theta / (1.0 - Theta);
          double thp;
          if (g_mirror) { thp = 1.0; for (int k = 0; k < tb.Nit - 1 - Newt; k++) thp = thp * Theta; }
          else thp = pow(Theta, (double)(tb.Nit - 1 - Newt));
          double dyth = FacConv * NewNrm * thp / FNewt;
          if (dyth >= 1.0) {
            double qNewt = fmax(1.0e-4, fmin(20.0, dyth));
            double hhfac = 0.8 * POW(qNewt, -1.0 / (4.0 + tb.Nit - 1 - Newt));
            h = hhfac * h;
            Reject = 1; Last = 0;
            UnExpNewtRej = hhfac <= 0.5;
            st[ST_NREJ] += 1; st[ST_REJ] += 1;
            NeedNewLU = 1; newton_failed = 1;
            break;
          }
        } else {
          h = h * 0.5;
          Reject = 1; Last = 0; UnExpStepRej = 1;
          st[ST_REJ] += 1;
          NeedNewLU = 1; newton_failed = 1;
          break;
        }
      }
      OldNrm = fmax(NewNrm, EPS);
      for (int q = 0; q < s; q++) for (int i = 0; i < n; i++) w[q][i] = w[q][i] + z[q][i];
      for (int j = 0; j < s; j++)
        for (int i = 0; i < n; i++) {
          double acc = tb.T[j][0] * w[0][i];
          for (int q = 1; q < s; q++) acc = FMA(tb.T[j][q], w[q][i], acc);
          z[j][i] = acc;
        }
      if (FacConv * NewNrm > FNewt) continue;
      break;
    }
    if (nan_fail) { res = 2; break; }
    if (g_plog && g_plog_n < g_plog_cap) {
      double *L = g_plog + 8 * g_plog_n++;
      L[0] = (double)st[ST_STEP]; L[1] = t; L[2] = h_attempt; L[3] = NbrStg; L[4] = Newt; L[5] = Theta; L[6] = -1.0; L[7] = FacConv;
    }
    if (newton_failed) continue;

    double err = estrad(pb, &tb, h, &lu, z, Scal, f0, First, Reject, t, y, st);
    if (g_plog && g_plog_n >= 1) g_plog[8 * (g_plog_n - 1) + 6] = err;
    if (!isfinite(err)) { res = 2; break; }
    double fac = fmin(Safe, (2.0 * tb.Nit + 1.0) / (2.0 * tb.Nit + Newt));
    double exponent = 1.0 / s + 1.0;
    double quot = fmax(FacR, fmin(FacL, POW(err, exponent) / fac));
    double hnew = h / quot;
    if (err < 1.0) {
      First = 0;
      st[ST_ACC] += 1;
      const int gus_upd = op->Gustafsson && !ChangeFlag;
      const int gus_lim = (g_quirks & Q8_GUS_UNGATED) ? op->Gustafsson : gus_upd;
      if (gus_lim) {
        if (st[ST_ACC] > 1) {
          double ratio = err * err / erracc, root;
          if (!g_mirror) root = pow(ratio, 1.0 / (s + 1.0));
          else if (s == 1) root = sqrt(ratio);
          else if (s == 3) root = sqrt(sqrt(ratio));
          else if (s == 7) root = sqrt(sqrt(sqrt(ratio)));
          else root = det_pow(ratio, 1.0 / (s + 1.0));
          double facgus = (hacc / h) * root / Safe;
          facgus = fmax(FacR, fmin(FacL, facgus));
          quot = fmax(quot, facgus);
          hnew = h / quot;
        }
      }
      if (gus_upd) {
        hacc = h;
        erracc = fmax(1.0e-2, err);
      }
      double E1v[4] = {0, 0, 0, 0};
      const double t_ev1 = t;
      if (dn && dn->ev_K > 0)
        for (int k = 0; k < dn->ev_K; k++) {
          double e = -dn->ev_b[k];
          for (int i = 0; i < n; i++) e = FMA(dn->ev_A[k * n + i], y[i], e);
          E1v[k] = e;
        }
      h_old = h;
      t = t + h;
      for (int i = 0; i < n; i++) y[i] = y[i] + z[s - 1][i];
      for (int i = 0; i < n; i++) cont[s - 1][i] = g_mirror ? z[0][i] * (1.0 / tb.C[0]) : z[0][i] / tb.C[0];
      for (int q = 1; q <= s - 1; q++) {
        double Fact = 1.0 / (tb.C[s - q - 1] - tb.C[s - q]);
        for (int i = 0; i < n; i++) cont[q - 1][i] = (z[s - q - 1][i] - z[s - q][i]) * Fact;
      }
      for (int jj = 2; jj <= s; jj++)
        for (int k = s; k >= jj; k--) {
          double Fact = (s - k == 0) ? 1.0 / (-tb.C[jj - 1]) : 1.0 / (tb.C[s - k - 1] - tb.C[s - k + jj - 1]);
          for (int i = 0; i < n; i++) cont[k - 1][i] = (cont[k - 1][i] - cont[k - 2][i]) * Fact;
        }
      if (rec_h && rec_n && *rec_n < rec_cap) rec_h[(*rec_n)++] = h;
      if (dn && dn->ev_K > 0 && locate_events(n, s, &tb, cont, dn, E1v, t_ev1, &t, y, h_old)) { finished = 1; break; }
      if (dn && (dn->mode == 1 || dn->mode == 2)) {
        for (int r = 1; r < dn->refine; r++) {
          const double tint = t + h_old * ((double)r / (double)dn->refine) - h_old;
          if (dn->nout < dn->cap) {
            dn->tout[dn->nout] = tint;
            ntrp(n, s, &tb, cont, t, y, h_old, tint, dn->yout + dn->nout * n);
          }
          dn->nout += 1;
        }
        if (dn->nout < dn->cap) {
          dn->tout[dn->nout] = t;
          for (int i = 0; i < n; i++) dn->yout[dn->nout * n + i] = y[i];
        }
        dn->nout += 1;
      } else if (dn && dn->tsp) { /* radau.m:1150-1168 + ntrprad :1291-1308 (h_old = the step just taken) */
        while (dn->nout < dn->ntspan) {
          const double tn = dn->tsp[dn->nout];
          if (PosNeg * (t - tn) < 0) break;
          if (t == tn) {
            dn->tout[dn->nout] = t;
            for (int i = 0; i < n; i++) dn->yout[dn->nout * n + i] = y[i];
            dn->nout += 1;
            break;
          }
          const double sfr = (tn - t) / h_old;
          for (int i = 0; i < n; i++) {
            double yi = (sfr - (tb.C[0] - 1.0)) * cont[s - 1][i];
            for (int q = 1; q < s; q++) yi = (sfr - (tb.C[q] - 1.0)) * (yi + cont[s - 1 - q][i]);
            dn->yout[dn->nout * n + i] = yi + y[i];
          }
          dn->tout[dn->nout] = tn;
          dn->nout += 1;
        }
      }
      NeedNewJac = 1;
      if (Last) {
        h = hopt;
        st[ST_REJ] += 1;
        finished = 1;
        break;
      }
      rhs(pb, t, y, f0);
      st[ST_FCN] += 1;
      if (any_nan(f0, n)) { res = 2; break; }
      hnew = PosNeg * fmin(fabs(hnew), fabs(hmaxn));
      hopt = PosNeg * fmin(fabs(h), fabs(hnew));
      if (Reject) hnew = PosNeg * fmin(fabs(hnew), fabs(h));
      Reject = 0;
      if ((t + hnew / Quot1 - tfinal) * PosNeg >= 0) {
        h = tfinal - t;
        Last = 1;
      } else {
        double qt = hnew / h;
        if (Theta <= Thet && qt >= Quot1 && qt <= Quot2) {
          Keep = 1; NeedNewJac = 0; NeedNewLU = 0;
          continue;
        }
        h = hnew;
      }
      NeedNewLU = 1;
      if (Theta <= Thet) NeedNewJac = 0;
    } else {
      Reject = 1; Last = 0;
      if (First) h = h / 10.0; else h = hnew;
      if (st[ST_ACC] >= 1) st[ST_REJ] += 1;
      NeedNewLU = 1;
    }
  }
  /* radau.m:1287: only the step cap ends the loop early (t + (tfinal - t) may round below tfinal) */
  if (res == 0 && !finished && PosNeg * t < PosNeg * tfinal) res = 3;
done:
  if (dn && (dn->mode == 1 || dn->mode == 2)) {
    if (dn->nout >= 1 && dn->nout <= dn->cap) {
      dn->tout[dn->nout - 1] = t;
      for (int i = 0; i < n; i++) dn->yout[(dn->nout - 1) * n + i] = y[i];
    }
  } else if (dn && dn->tsp && dn->nout >= 1) { /* radau.m:1267-1269 */
    dn->tout[dn->nout - 1] = t;
    for (int i = 0; i < n; i++) dn->yout[(dn->nout - 1) * n + i] = y[i];
  }
  for (int i = 0; i < n; i++) y_out[i] = y[i];
  *t_out = t;
  *h_out = h;
  *nstg_out = NbrStg;
  return res;
}

/* ------------------------------------------------------------------ public entry points */
/* y0 [M][Ny] row-major, tspan [M][2]; outputs y_out [M][Ny], t_out[M], h_out[M], res[M],
 * stats [M][8], nstg[M].  nthreads <= 0 -> all.  Returns 0.                               */
void oracle_set_quirks(int q) { g_quirks = q; }
int oracle_get_quirks(void) { return g_quirks; }
void oracle_set_arith(int mirror) { g_mirror = mirror ? 1 : 0; }
int oracle_get_arith(void) { return g_mirror; }

int oracle_radau_solve(int problem, const double *par, int sign, int Ny, long long M, const double *y0,
                       const double *tspan, const oracle_opts *op, double *y_out, double *t_out, double *h_out,
                       int *res, long long *stats, int *nstg, int nthreads) {
  problem_t pb;
  pb.id = problem; pb.Ny = Ny; pb.sign = (double)sign; pb.par = par;
  if (Ny > NYMAX) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (long long m = 0; m < M; m++) {
    res[m] = radau_one(&pb, op, tspan[2 * m], tspan[2 * m + 1], y0 + m * Ny, y_out + m * Ny, t_out + m,
                       h_out + m, stats + 8 * m, nstg + m, NULL, 0, NULL, NULL);
  }
  return 0;
}

/* same with output at the interior points of tspan [M][ntspan]: tout [M][ntspan], yout [M][ntspan][Ny], nout [M] */
int oracle_radau_solve_dense(int problem, const double *par, int sign, int Ny, long long M, const double *y0,
                             const double *tspan, int ntspan, const oracle_opts *op, double *y_out, double *t_out,
                             double *h_out, int *res, long long *stats, int *nstg, double *tout, double *yout,
                             int *nout, int nthreads) {
  problem_t pb;
  pb.id = problem; pb.Ny = Ny; pb.sign = (double)sign; pb.par = par;
  if (Ny > NYMAX) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (long long m = 0; m < M; m++) {
    dense_t dn;
    memset(&dn, 0, sizeof dn);
    dn.tsp = tspan + m * ntspan; dn.ntspan = ntspan; dn.tout = tout + m * ntspan;
    dn.yout = yout + m * ntspan * Ny; dn.nout = 0;
    res[m] = radau_one(&pb, op, tspan[m * ntspan], tspan[m * ntspan + ntspan - 1], y0 + m * Ny, y_out + m * Ny,
                       t_out + m, h_out + m, stats + 8 * m, nstg + m, NULL, 0, NULL, &dn);
    nout[m] = dn.nout;
  }
  return 0;
}

/* output at every accepted step (OutFlag 1 / 2) and located events; tspan [M][2] */
int oracle_radau_solve_out(int problem, const double *par, int sign, int Ny, long long M, const double *y0,
                           const double *tspan, const oracle_opts *op, double *y_out, double *t_out, double *h_out, int *res,
                           long long *stats, int *nstg, int refine, int cap, double *tout, double *yout, int *nout,
                           int ev_K, const double *ev_A, const double *ev_b, const int *ev_stop, const int *ev_slope,
                           int ecap, double *te, double *ye, int *ie, int *nev, int nthreads) {
  problem_t pb;
  pb.id = problem; pb.Ny = Ny; pb.sign = (double)sign; pb.par = par;
  if (Ny > NYMAX) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (long long m = 0; m < M; m++) {
    dense_t dn;
    memset(&dn, 0, sizeof dn);
    dn.mode = refine > 1 ? 2 : 1; dn.cap = cap; dn.refine = refine;
    dn.tout = tout + m * cap; dn.yout = yout + m * cap * Ny;
    dn.ev_K = ev_K; dn.ev_A = ev_A; dn.ev_b = ev_b; dn.ev_stop = ev_stop; dn.ev_slope = ev_slope;
    dn.te = te + m * ecap; dn.ye = ye + m * ecap * Ny; dn.ie = ie + m * ecap; dn.ecap = ecap;
    res[m] = radau_one(&pb, op, tspan[2 * m], tspan[2 * m + 1], y0 + m * Ny, y_out + m * Ny, t_out + m, h_out + m,
                       stats + 8 * m, nstg + m, NULL, 0, NULL, &dn);
    nout[m] = dn.nout;
    nev[m] = dn.nev;
  }
  return 0;
}

/* single trajectory with the accepted step sizes recorded (debug / desync analysis) */
int oracle_radau_trace(int problem, const double *par, int sign, int Ny, const double *y0, const double *tspan,
                       const oracle_opts *op, double *y_out, double *t_out, long long *stats, double *rec_h,
                       int rec_cap, int *rec_n) {
  problem_t pb;
  pb.id = problem; pb.Ny = Ny; pb.sign = (double)sign; pb.par = par;
  double h_out; int nstg;
  return radau_one(&pb, op, tspan[0], tspan[1], y0, y_out, t_out, &h_out, stats, &nstg, rec_h, rec_cap, rec_n, NULL);
}

/* ------------------------------------------------------------------ SEULEX (semi-implicit extrapolation)
 * One call = one step() of the reference's Python stepper, per trajectory:
 *   /root/reference/src/python/seulext.py:8-64 (SeulexState: nseq, cost, coeff, k_targ, theta, jac_redo),
 *   :158-581 (SeulexT.step: order / step-size control), :593-667 (dy_batch: the linearly implicit Euler sequence
 *   with the stability check), polyextr (src/python/seulex.py:439-462).  The reference's masked code is restated
 *   for a batch of ONE trajectory (M = 1), which is how it is run live for the golden vectors
 *   (tests/golden/make_seulex_golden.py).  Its algorithm is Numerical Recipes' StepperSie; the reference's
 *   deviations from it are kept (they are what the live stepper does):
 *     S2  after a rejected attempt without a fresh Jacobian: theta *= jac_redo (seulext.py:476-478; NR: theta =
 *         2 jac_redo and the Jacobian is re-evaluated)
 *     S3  kopt is capped at KMAXX, not KMAXX - 1 (:511, :527)
 *     S4  in the k == k_targ - 1 branch the new step size / break only happen when k_targ is also lowered (:378-389)
 *     S5  the sequence is integrated with |h| (forward in time only, :249)
 *   and one is switchable, default OFF because it corrupts the time axis:
 *     S1  quirks & 1: the clock advances by htry even when the step was shortened after a rejected attempt
 *         (x += h with h = htry, :483-485: `h` is never updated inside compute_jac)
 * flags: 1 first_step, 2 last_step, 4 reject, 8 prev_reject.                                                   */
#define SX_IMAXX 16
static void seulex_one(const problem_t *pb, double atol, double rtol, int KMAXX, int quirks, double htry, double *y,
                       double *x_io, double *theta_io, int *ktarg_io, double *hnext_io, double *hdid_out, int *flags_io,
                       double *dfdy, int *k_out, int *nrej_out) {
  const int n = pb->Ny, IMAXX = KMAXX + 1;
  int nseq[SX_IMAXX + 1];
  double cost[SX_IMAXX + 1], coeff[SX_IMAXX + 1][SX_IMAXX + 1], hopt[SX_IMAXX + 1], work[SX_IMAXX + 1];
  double table[SX_IMAXX][NYMAX], ysav[NYMAX], yseq[NYMAX], scale[NYMAX];
  nseq[0] = 2; nseq[1] = 3;
  for (int i = 2; i < IMAXX; i++) nseq[i] = 2 * nseq[i - 2];
  cost[0] = 5.0 + 1.0 + nseq[0] * (1.0 + 1.0);
  for (int k = 0; k < KMAXX; k++) cost[k + 1] = cost[k] + (nseq[k + 1] - 1) * (1.0 + 1.0) + 1.0;
  for (int k = 0; k < IMAXX; k++)
    for (int l = 0; l < k; l++) coeff[k][l] = 1.0 / ((double)nseq[k] / (double)nseq[l] - 1.0);
  const double jac_redo = fmin(1.0e-4, rtol);
  double x = *x_io, theta = *theta_io, hnext = *hnext_io;
  int k_targ = *ktarg_io, fl = *flags_io;
  int first_step = fl & 1, last_step = (fl >> 1) & 1, reject = (fl >> 2) & 1, prev_reject = (fl >> 3) & 1, calcjac = 0;
  double errold = 0.0, fac = 0.0;
  for (int i = 0; i < IMAXX; i++) { hopt[i] = 0.0; work[i] = 0.0; }
  work[0] = 1.0e30;
  const double h = htry;
  for (int i = 0; i < n; i++) ysav[i] = y[i];
  const int forward = h > 0;
  if (h != hnext && !first_step) last_step = 1;
  if (reject) { last_step = 0; prev_reject = 1; theta = 2.0 * jac_redo; }
  for (int i = 0; i < n; i++) scale[i] = FMA(rtol, fabs(y[i]), atol);
  reject = 0;
  double hnew = fabs(h), h_att = fabs(h);
  int k = 0, firstk = 1, nrej = 0;
  while (firstk || reject) {
    if (theta > jac_redo && !calcjac && fabs(h) > EPS) {
      jacobian(pb, x, y, dfdy);
      calcjac = 1;
    }
    firstk = 0; reject = 0;
    h_att = hnew;
    int brk = 0;
    k = 0;
    while (k <= k_targ + 1 && !brk) {
      /* ---- dy: nseq[k] linearly implicit Euler steps over h_att (seulext.py:593-667) */
      int success = 1;
      {
        const int nsteps = nseq[k];
        const double hh = hnew / nsteps, rhh = 1.0 / hh;
        rlu_t lu;
        for (int i = 0; i < n; i++)
          for (int j = 0; j < n; j++) lu.a[i * n + j] = (i == j) ? (-dfdy[i * n + j] + rhh) : -dfdy[i * n + j];
        rlu_factor(n, &lu);
        double xnew = x + hh, ytemp[NYMAX], delta[NYMAX], yend[NYMAX];
        for (int i = 0; i < n; i++) ytemp[i] = ysav[i];
        rhs(pb, xnew, ytemp, delta);
        rlu_solve(n, &lu, delta);
        for (int nn = 1; nn < nsteps; nn++) {
          for (int i = 0; i < n; i++) ytemp[i] = ytemp[i] + delta[i];
          xnew += hh;
          rhs(pb, xnew, ytemp, yend);
          if (nn == 1 && k <= 1) {
            double del1 = 0.0, del2 = 0.0, dyt[NYMAX];
            for (int i = 0; i < n; i++) { const double v = delta[i] / scale[i]; del1 = FMA(v, v, del1); }
            del1 = sqrt(del1);
            rhs(pb, x + hh, ytemp, dyt);
            for (int i = 0; i < n; i++) delta[i] = dyt[i] - delta[i] * rhh;
            rlu_solve(n, &lu, delta);
            for (int i = 0; i < n; i++) { const double v = delta[i] / scale[i]; del2 = FMA(v, v, del2); }
            del2 = sqrt(del2);
            theta = del2 / fmax(1.0, del1);
            if (theta > 1.0) { success = 0; break; }
          }
          rlu_solve(n, &lu, yend);
          for (int i = 0; i < n; i++) delta[i] = yend[i];
        }
        for (int i = 0; i < n; i++) yseq[i] = ytemp[i] + delta[i];
      }
      if (!success) { hnew = fabs(hnew) * 0.5; reject = 1; brk = 1; break; }
      if (k == 0) {
        for (int i = 0; i < n; i++) y[i] = yseq[i];
      } else {
        for (int i = 0; i < n; i++) table[k - 1][i] = yseq[i];
        for (int j = k - 1; j > 0; j--)   /* polyextr */
          for (int i = 0; i < n; i++) table[j - 1][i] = FMA(coeff[k][j], table[j][i] - table[j - 1][i], table[j][i]);
        for (int i = 0; i < n; i++) y[i] = FMA(coeff[k][0], table[0][i] - y[i], table[0][i]);
        double err = 0.0;
        for (int i = 0; i < n; i++) {
          scale[i] = FMA(rtol, fabs(ysav[i]), atol);
          const double v = (y[i] - table[0][i]) / scale[i];
          err = FMA(v, v, err);
        }
        err = sqrt(err / n);
        if (err > 1.0 / EPS || (k > 1 && err >= errold)) { reject = 1; hnew = fabs(hnew) * 0.5; brk = 1; break; }
        errold = fmax(4.0 * err, 1.0);
        const double expo = 1.0 / (k + 1);
        const double facmin = POW(0.1, expo);
        if (err == 0.0) fac = 1.0 / facmin;
        else {
          fac = 0.93 / POW(err / 0.6, expo);
          fac = fmax(facmin / 4.0, fmin(1.0 / facmin, fac));
        }
        hopt[k] = fabs(hnew * fac);
        work[k] = cost[k] / hopt[k];
        if ((first_step || last_step) && err <= 1.0) { brk = 1; break; }
        if (k == k_targ - 1 && !prev_reject && !first_step && !last_step) {
          if (err <= 1.0) { brk = 1; break; }
          if (err > (double)nseq[k_targ] * (double)nseq[k_targ + 1] * 4.0) {
            reject = 1; k_targ = k;
            if (k_targ > 1 && work[k - 1] < 0.7 * work[k]) {   /* S4: step size and break only in here */
              k_targ -= 1;
              hnew = hopt[k_targ];
              brk = 1; break;
            }
          }
        }
        if (k == k_targ) {
          if (err <= 1.0) { brk = 1; break; }
          if (err > nseq[k + 1] * 2.0) {
            reject = 1;
            if (k_targ > 1 && work[k - 1] < 0.7 * work[k]) k_targ -= 1;
            hnew = hopt[k_targ];
            brk = 1; break;
          }
        }
        if (k == k_targ + 1) {
          if (err > 1.0) {
            reject = 1;
            if (k_targ > 1 && work[k_targ - 1] < 0.7 * work[k_targ]) k_targ -= 1;
            hnew = hopt[k_targ];
          }
          brk = 1; break;
        }
      }
      k += 1;
    }
    if (reject) {
      nrej += 1;
      prev_reject = 1;
      if (!calcjac) theta = theta * jac_redo;   /* S2 */
      if (nrej > 200) break;                    /* guard of this restatement only: a step that can never succeed */
    }
  }
  const double hs = (quirks & 1) ? h : (forward ? h_att : -h_att);   /* S1 */
  x = x + hs;
  *hdid_out = hs;
  first_step = 0;
  int kopt = k;
  if (k == 1) kopt = 2;
  if (k <= k_targ) {
    kopt = k;
    if (work[k - 1] < work[k] * 0.7) kopt = k - 1;
    if (work[k] < 0.9 * work[k - 1]) kopt = imin(k + 1, KMAXX);
  }
  if (k != 1 && k > k_targ) {
    kopt = k - 1;
    if (k > 2 && work[k - 2] < 0.7 * work[k - 1]) kopt = k - 2;
    if (work[k] < 0.9 * work[kopt]) kopt = imin(k, KMAXX);
  }
  if (prev_reject) {
    k_targ = imin(kopt, k);
    hnew = fmin(fabs(hs), hopt[k_targ]);
    prev_reject = 0;
  } else {
    if (kopt <= k) hnew = hopt[kopt];
    else if (k < k_targ && work[k] < 0.9 * work[k - 1]) hnew = hopt[k] * cost[imin(kopt + 1, KMAXX)] / cost[k];
    else hnew = hopt[k] * cost[kopt] / cost[k];
    k_targ = kopt;
  }
  *hnext_io = forward ? hnew : -hnew;
  *x_io = x; *theta_io = theta; *ktarg_io = k_targ;
  *flags_io = (first_step ? 1 : 0) | (last_step ? 2 : 0) | (reject ? 4 : 0) | (prev_reject ? 8 : 0);
  *k_out = k; *nrej_out = nrej;
}

/* initial controller state of SeulexState (seulext.py:44-46, 66-70) */
void oracle_seulex_init(double atol, double rtol, int KMAXX, double *theta, int *k_targ) {
  const double jac_redo = fmin(1.0e-4, rtol);
  *theta = 2.0 * jac_redo;
  const double logfact = -log10(rtol + atol) * 0.6 + 0.5;
  double kt = fmax(1.0, fmin((double)(KMAXX - 1), nearbyint(logfact)));
  *k_targ = (int)floor(kt);
}

int oracle_seulex_step(int problem, const double *par, int sign, int Ny, long long M, double atol, double rtol, int KMAXX,
                       int quirks, const double *htry, double *y, double *x, double *theta, int *k_targ, double *hnext,
                       double *hdid, int *flags, double *dfdy, int *k_out, int *nrej) {
  if (KMAXX + 1 > SX_IMAXX) return -1;
  problem_t pb;
  pb.id = problem; pb.Ny = Ny; pb.sign = (double)sign; pb.par = par;
  for (long long m = 0; m < M; m++)
    seulex_one(&pb, atol, rtol, KMAXX, quirks, htry[m], y + m * Ny, x + m, theta + m, k_targ + m, hnext + m, hdid + m,
               flags + m, dfdy + m * Ny * Ny, k_out + m, nrej + m);
  return 0;
}

/* batched dynamics / Jacobian (checks the stage-eval kernel) */
int oracle_radau_trace2(int problem, const double *par, int sign, int Ny, const double *y0, const double *tspan,
                        const oracle_opts *op, double *y_out, long long *stats, double *plog, int cap, int *n) {
  g_plog = plog; g_plog_cap = cap; g_plog_n = 0;
  double t; int rn = 0;
  int rc = oracle_radau_trace(problem, par, sign, Ny, y0, tspan, op, y_out, &t, stats, 0, 0, &rn);
  *n = g_plog_n; g_plog = 0;
  return rc;
}

int oracle_eval_dyns(int problem, const double *par, int sign, int Ny, long long M, const double *y, double *f) {
  problem_t pb; pb.id = problem; pb.Ny = Ny; pb.sign = (double)sign; pb.par = par;
#pragma omp parallel for
  for (long long m = 0; m < M; m++) rhs(&pb, 0.0, y + m * Ny, f + m * Ny);
  return 0;
}

int oracle_eval_jac(int problem, const double *par, int sign, int Ny, long long M, const double *y, double *J) {
  problem_t pb; pb.id = problem; pb.Ny = Ny; pb.sign = (double)sign; pb.par = par;
#pragma omp parallel for
  for (long long m = 0; m < M; m++) jacobian(&pb, 0.0, y + m * Ny, J + m * Ny * Ny);
  return 0;
}

/* batched LU factor+solve: A [M][n][n] (complex as separate re/im, im may be NULL), b [M][n] */
int oracle_lu_solve(int n, long long M, const double *Are, const double *Aim, const double *bre, const double *bim,
                    double *xre, double *xim) {
  if (n > NYMAX) return -1;
#pragma omp parallel for
  for (long long m = 0; m < M; m++) {
    if (Aim == NULL) {
      rlu_t L;
      memcpy(L.a, Are + m * n * n, sizeof(double) * n * n);
      rlu_factor(n, &L);
      double b[NYMAX];
      memcpy(b, bre + m * n, sizeof(double) * n);
      rlu_solve(n, &L, b);
      memcpy(xre + m * n, b, sizeof(double) * n);
    } else {
      zlu_t L;
      memcpy(L.re, Are + m * n * n, sizeof(double) * n * n);
      memcpy(L.im, Aim + m * n * n, sizeof(double) * n * n);
      zlu_factor(n, &L);
      double br[NYMAX], bi[NYMAX];
      memcpy(br, bre + m * n, sizeof(double) * n);
      memcpy(bi, bim + m * n, sizeof(double) * n);
      zlu_solve(n, &L, br, bi);
      memcpy(xre + m * n, br, sizeof(double) * n);
      memcpy(xim + m * n, bi, sizeof(double) * n);
    }
  }
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* deterministic pow (mirror of csrc/detmath.cuh) exposed for the bitwise test */
void oracle_det_pow(const double *x, const double *y, long long n, double *out) {
  for (long long i = 0; i < n; i++) out[i] = det_pow(x[i], y[i]);
}
void oracle_det_tanh(const double *x, long long n, double *out) {
  for (long long i = 0; i < n; i++) out[i] = det_tanh(x[i]);
}
