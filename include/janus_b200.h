/*
 * janus_b200.h -- C ABI of the B200-native batched Hamiltonian Radau integrator.
 *
 * Drop-in boundary for ONE path of optstat/janus-ode: the batched implicit Radau IIA solve of
 * a Hamiltonian system (H -> dynamics/Jacobian -> simplified Newton + LU -> error estimate ->
 * per-trajectory step/order control).  Plain pointers and sizes only, no torch types.
 * Every function returns 0 on success or a negative JB200_E* code; jb200_last_error() gives
 * the text.  Nothing here ever calls exit() (the reference does: radaute_impl.hpp:10-214).
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   jb200_opts              <- struct OptionsTe               src/cpp/radaute.hpp:52-96
 *   jb200_create            <- RadauTe::RadauTe(OdeFcn, JacFn, tspan, y0, Op, params)
 *                                                             src/cpp/radaute.hpp:424-425,
 *                                                             src/cpp/radaute_impl.hpp:5-535
 *                              with OdeFcn/JacFn = evalDyns/evalJac of a Hamiltonian
 *                                                             src/cpp/janus_ode_common.hpp:2161-2299
 *                                                             examples/cpp/radaute_hamiltonian_example.cpp:44-62
 *   jb200_solve[_host]      <- Tensor RadauTe::solve() + public fields y, t, h, stats
 *                                                             src/cpp/radaute.hpp:427,296,253,98-108
 *                                                             src/cpp/radaute_impl.hpp:537-1345
 *   jb200_eval_dyns/_jac    <- evalDyns(U) / evalJac(U)       src/cpp/janus_ode_common.hpp:2161-2299
 *   jb200_lu_solve          <- torch::linalg_lu_factor + lu_solve on [m,Ny,Ny] (the janus-linear
 *                              role)                          src/cpp/radaute_impl.hpp:1558,1611,1754,1832
 * Layouts are the reference's: y [M,Ny] row-major with y = [p(0:N), x(N:2N)]
 * (janus_ode_common.hpp:2167-2169), tspan [M,2], J [M,Ny,Ny] rows = ydot, cols = y.
 *
 * Thread-compatibility: one handle per (device, stream); calls on one handle must not overlap.
 */
#ifndef JANUS_B200_H
#define JANUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB200_NYMAX 16
#define JB200_MAX_HAM_PARAMS 8

/* error codes */
#define JB200_OK 0
#define JB200_EINVAL (-1)   /* bad option / argument (the reference exit(1)s) */
#define JB200_ECUDA (-2)    /* CUDA runtime error                              */
#define JB200_ENOTSUP (-3)  /* Hamiltonian / size not compiled in              */

/* built-in Hamiltonians, y = [p, x] */
#define JB200_HAM_HARMONIC 0        /* N=1 par: omega              H=(p^2+w^2 x^2)/2                      */
#define JB200_HAM_VDP 1             /* N=2 par: mu                 H=p1 x2+p2(mu(1-x1^2)x2-x1)            */
#define JB200_HAM_VDP_PMP_SMOOTH 2  /* N=2 par: W                  Hu=p1x2+p2u((1-x1^2)x2-x1)+W/2u^2+1,
                                                                   u=-p2((1-x1^2)x2-x1)/W frozen         */
#define JB200_HAM_VDP_PMP_BANG 3    /* N=2 par: W,umin,umax        same Hu, u=umax if p2(1-x1^2)x2<0 else umin */
#define JB200_HAM_VDP_AUG_BANG 4    /* N=2 par: umin,umax          H=p1x2+p2(u(1-x1^2)x2-x1), bang-bang u */
#define JB200_HAM_MLP 5             /* N=8 par: [8, 128, W1[128][16], b1[128], W2[128][128], b2, W3[128][128], b3,
                                       W4[128], b4] (35,331 numbers)  H = W4.tanh(W3 tanh(W2 tanh(W1 [x;p]+b1)+b2)+b3)+b4:
                                       stage evaluations as batched fp64 GEMM tiles (csrc/mlp_kernels.cuh)         */
#define JB200_HAM_USER 100          /* plugin libraries only: the user-supplied Hamiltonian the plugin was
                                       compiled around (janus_ode_b200/user.py; same ABI, N <= 2)          */

/* result codes per trajectory (res[]) -- src/cpp/radaute_impl.hpp:711-719,873-877,1245-1255 */
#define JB200_RES_OK 0
#define JB200_RES_STEP_TOO_SMALL 1
#define JB200_RES_NAN 2
#define JB200_RES_MAX_STEPS 3
#define JB200_RES_SENS_UNCONVERGED 4   /* jb200_solve_sens only: the state converged, a sensitivity iteration did not */

/* stats columns: StatsTe, src/cpp/radaute.hpp:98-108 (MATLAB counting rules, radau.m) */
enum {
  JB200_ST_FCN = 0, JB200_ST_JAC, JB200_ST_DECOMP, JB200_ST_SOLVE,
  JB200_ST_STEP, JB200_ST_ACCPT, JB200_ST_STEPREJ, JB200_ST_NEWTREJ, JB200_NSTATS
};

/* mirrors OptionsTe field for field (solver knobs), plus costate_sign */
typedef struct jb200_opts {
  double RelTol[JB200_NYMAX];   /* n_tol entries used (1 = scalar, broadcast) */
  double AbsTol[JB200_NYMAX];
  int n_tol;
  double InitialStep;           /* 1e-2 */
  double MaxStep;               /* accepted and ignored like the reference (hmax = tfinal - t0) */
  double JacRecompute;          /* 1e-3 */
  int Start_Newt;               /* 0 */
  int MaxNbrNewton;             /* 7 */
  int NbrStg, MinNbrStg, MaxNbrStg; /* 3, 3, 7 */
  int64_t MaxNbrStep;           /* 10000 */
  double Safe, Quot1, Quot2;    /* 0.9, 1, 1.2 */
  double FacL, FacR;            /* 0.2, 8 */
  double Vitu, Vitd, hhou, hhod;/* 0.002, 0.8, 1.2, 0.8 */
  int Gustafsson;               /* 1 */
  int costate_sign;             /* +1: pdot=+dH/dx (reference), -1: canonical */
  int radaute_quirks;           /* bit mask of RadauTe deviations from its MATLAB twin to reproduce (default: all):
                                   JB200_Q8_GUSTAFSSON_UNGATED -- the Gustafsson limiter also acts on accepted steps whose
                                   order just changed (src/cpp/radaute_impl.hpp:1092-1100 vs src/matlab/radau.m:1066-1076) */
} jb200_opts;
#define JB200_Q8_GUSTAFSSON_UNGATED 1

/* fill *o with the reference defaults (src/cpp/radaute.hpp:52-96) */
int jb200_opts_default(jb200_opts* o);

/* validate options + Hamiltonian and build a solver handle on the current CUDA device */
int jb200_create(const jb200_opts* o, int ham_id, const double* ham_params, int n_ham_params,
                 void** handle);
void jb200_destroy(void* handle);

/* number of state components Ny = 2N of the handle's Hamiltonian */
int jb200_ny(void* handle);

/* Integrate M trajectories.  All pointers are DEVICE pointers; stream is a cudaStream_t.
 * y0 [M,Ny], tspan [M,2] -> y_out [M,Ny], t_out [M], h_out [M], res [M], stats [M,8] (int64),
 * nstg [M] (final stage count; may be NULL), work [M] (algorithmic FLOPs per trajectory by the
 * SURVEY 8d model, for roofline accounting; may be NULL).  Asynchronous on `stream`.       */
int jb200_solve(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out,
                double* t_out, double* h_out, int32_t* res, int64_t* stats, int32_t* nstg,
                int64_t* work, void* stream);

/* Output at interior points of tspan (OutFlag 3 of the reference's MATLAB twin,
 * src/matlab/radau.m:1150-1168, ntrprad :1291-1308 == src/cpp/radaute_impl.hpp:1413-1440; the
 * C++ RadauTe leaves this case as a TODO, :1206-1207).  tspan [M,ntspan] with ntspan >= 2 (first
 * column t0, last tfinal); after every accepted step the requested times it covers are evaluated
 * on the collocation polynomial: tout [M,ntspan], yout [M,ntspan,Ny], nout [M] (points written;
 * the last written point is overwritten with the final (t, y) like radau.m:1267-1269).  Rows
 * beyond nout are left untouched.  Device pointers.                                        */
int jb200_solve_dense(void* handle, const double* y0, const double* tspan, int ntspan, int64_t M,
                      double* y_out, double* t_out, double* h_out, int32_t* res, int64_t* stats,
                      int32_t* nstg, int64_t* work, double* tout, double* yout, int32_t* nout,
                      void* stream);

/* Output at every accepted step and event location.
 * jb200_set_events: event functions of the handle, E_k(t, y) = A[k,:] . y - b[k], k < K <= 4, with the reference's
 * `isterminal` (stop[k]) and `direction` (slope[k] = +1 rising, -1 falling, 0 both) -- EventFnType,
 * src/cpp/radaute.hpp:44-47; zero location = EventZeroFcn, src/matlab/radau.m:1494-1618 (regula falsi / pegasus on the
 * collocation polynomial; the C++ port src/cpp/radaute_impl.hpp:2030-2185 is incomplete).  K = 0 removes them.  A kernel
 * cannot call a Python / std::function callback, hence the closed family of linear event functions.
 * jb200_solve_out: tspan [M,2]; the computed points are recorded like the reference's OutFlag 1 (refine = 1,
 * radaute_impl.hpp:1166-1183) / OutFlag 2 (refine > 1: refine - 1 interpolated points before each, radau.m:1130-1143) into
 * tout [M,cap], yout [M,cap,Ny]; nout [M] counts ALL points (nout > cap: the caller's buffer was too small -- the
 * reference regrows its buffer instead); the last stored point is overwritten with the final (t, y) like radau.m:1267-1269.
 * Events (if set): te [M,ecap], ye [M,ecap,Ny], ie [M,ecap] (1-based index of the event function), nev [M]; a terminal
 * event ends the trajectory at (te, ye) with res 0.  cap = 0 / ecap = 0: count only.  Device pointers.             */
int jb200_set_events(void* handle, int K, const double* A, const double* b, const int32_t* stop, const int32_t* slope);
int jb200_solve_out(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out, double* t_out,
                    double* h_out, int32_t* res, int64_t* stats, int32_t* nstg, int64_t* work, int refine, int cap,
                    double* tout, double* yout, int32_t* nout, int ecap, double* te, double* ye, int32_t* ie,
                    int32_t* nev, void* stream);

/* Same with HOST pointers (what a reference caller holds: CPU tensors).  Copies in, solves,
 * copies out and synchronises; the device workspace lives in the handle.                   */
int jb200_solve_host(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out,
                     double* t_out, double* h_out, int32_t* res, int64_t* stats, int32_t* nstg,
                     int64_t* work);

/* evalDynsDual / evalJacDual (+ U twins) and the eight third-order ladders ppppppH ... pxpxpxH of the reference
 * (src/cpp/janus_ode_common.hpp:2075-2158, :2302-2416, :787-2072) from ONE third-order jet of H per trajectory.
 * y [M,Ny] -> f [M,Ny], J [M,Ny,Ny] (as jb200_eval_dyns / _jac), T [M,Ny,Ny,Ny] with T[r][c][k] = d J[r][c] / d y_k
 * (the third-order blocks in y-order, costate sign on the pdot rows; may be NULL); with ND dual directions
 * yd [M,Ny,ND] (the `.d` of the reference's TensorDual y): fd [M,Ny,ND] = J yd, Jd [M,Ny,Ny,ND] = T yd (the `.d` parts
 * of the returned TensorDual / TensorMatDual; may be NULL).  Device pointers.                                   */
int jb200_eval_jac_dual(void* handle, const double* y, const double* yd, int ND, int64_t M, double* f, double* J,
                        double* T, double* fd, double* Jd, void* stream);

/* SEULEX (semi-implicit extrapolation) behind the same boundary: one call = one `SeulexT.step(htry, dyns, state)` of the
 * reference's Python stepper for every trajectory (src/python/seulext.py:158-581, dy_batch :593-667; state fields
 * of SeulexState :8-70).  Device pointers, updated in place: y [M,Ny], x [M] (time), theta [M], k_targ [M], hnext [M],
 * flags [M] (bit 0 first_step, 1 last_step, 2 reject, 3 prev_reject), yjac [M,Ny] (the state the kept Jacobian
 * `state.dfdy` was evaluated at); outputs hdid [M], k_out / nrej [M] (table rows used, rejected attempts; may be NULL).
 * quirks bit 0: the clock advances by htry even after an in-step reduction, like the reference (seulext.py:483-485);
 * 0 = advance by the step actually taken.  jb200_seulex_init gives the initial theta / k_targ of a SeulexState.   */
int jb200_seulex_step(void* handle, int64_t M, double atol, double rtol, int KMAXX, int quirks, const double* htry,
                      double* y, double* x, double* theta, int32_t* k_targ, double* hnext, double* hdid,
                      int32_t* flags, double* yjac, int32_t* k_out, int32_t* nrej, void* stream);

/* The same step with a dual part: replaces SeulexTed.step / SeulexTedState (/root/reference/src/python/seulexted.py:12-95,
 * 238-1396: TensorDual y, TensorMatDual dfdy, LUDual solves).  yd [M,Ny,nd] (1 <= nd <= 8) = the dual part of state.y, updated
 * in place; sjac [M,Ny,nd] = the dual part of the state the kept Jacobian was evaluated at (the reference keeps dfdy.d
 * itself), initialise to yd.  The accepted attempt is replayed on [y, y']: (I/hh - A) delta' = J(y_n) y_n' + A' delta with
 * A' = dJ/dy . sjac from the third-order jet of H; the extrapolation is linear.  Real part = jb200_seulex_step bit for bit. */
int jb200_seulex_step_dual(void* handle, int64_t M, double atol, double rtol, int KMAXX, int quirks, const double* htry, double* y,
                           double* x, double* theta, int32_t* k_targ, double* hnext, double* hdid, int32_t* flags, double* yjac,
                           int32_t* k_out, int32_t* nrej, double* yd, double* sjac, int32_t nd, void* stream);
int jb200_seulex_init(double atol, double rtol, int KMAXX, double* theta0, int32_t* k_targ0);

/* jb200_solve_host with a compact result set for the PCIe link: counters as int32 [M,8] (they are bounded by
 * MaxNbrStep * 17 * 7 < 2^31 for MaxNbrStep <= 1e6), and t_out / h_out / nstg / work optional (NULL = not copied):
 * 68 B per trajectory device->host for Ny = 4 instead of 128.  Same reference interface as jb200_solve_host
 * (RadauTe::solve() + the public result fields, src/cpp/radaute.hpp:296,253,98-108).                       */
int jb200_solve_host_compact(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out,
                             double* t_out, double* h_out, int32_t* res, int32_t* stats32, int32_t* nstg,
                             int64_t* work);

/* Incremental stepping: the state / step(htry, dyns, ft, state) protocol of the reference's Python
 * steppers (src/python/seulext.py:8-148 SeulexState, :158-581 step; src/python/seulex_dual_example_vdp.py:
 * 136-175 driver loop), carried by the Radau stepper.  A state holds M trajectories parked on the
 * device.  jb200_state_step advances every unfinished trajectory by exactly ONE ACCEPTED step
 * (rejected attempts are retried inside); htry [M] (device, may be NULL = keep the controller's own
 * proposal) is the trial step, as in step(htry, ...).  Stepping with htry = hnext until every
 * trajectory is done reproduces jb200_solve bit for bit.  jb200_state_read copies out y [M,Ny],
 * t [M], hnext [M], hdid [M], res [M], stats [M,8], nstg [M] (may be NULL), done [M] (may be NULL).
 * The handle must outlive its states.                                                        */
int jb200_state_create(void* handle, const double* y0, const double* tspan, int64_t M, void** state,
                       void* stream);
int jb200_state_step(void* state, const double* htry, void* stream);
int jb200_state_read(void* state, double* y, double* t, double* hnext, double* hdid, int32_t* res,
                     int64_t* stats, int32_t* nstg, int32_t* done, void* stream);
void jb200_state_destroy(void* state);

/* kernels launched by the last jb200_solve* call on this handle */
int jb200_last_launches(void* handle);

/* stage evaluation: ydot [M,Ny] and J [M,Ny,Ny] of the handle's Hamiltonian (device pointers) */
int jb200_eval_dyns(void* handle, const double* y, int64_t M, double* ydot, void* stream);
int jb200_eval_jac(void* handle, const double* y, int64_t M, double* jac, void* stream);

/* Solve with forward sensitivities: replaces RadauTeD(OdeFcnDual, JacFnDual, tspan, y0 (TensorDual), Op, params).solve()
 * (/root/reference/src/cpp/radauted.hpp, radauted_impl.hpp:620-1645) for Hamiltonian dynamics, the dual part of y0
 * being S0 [M,Ny,Ny] (NULL: identity = d y / d y0).  S_out [M,Ny,Ny] = the dual part of the final state (row i,
 * column d = d y_i(tfinal) / d direction d); sens_iters [M] (may be NULL) = sensitivity Newton iterations spent.
 * States, step sizes and counters are those of jb200_solve, bit for bit (step control looks at the real part only). */
int jb200_solve_sens(void* handle, const double* y0, const double* tspan, const double* S0, int64_t M, double* y_out,
                     double* S_out, double* t_out, double* h_out, int32_t* res, int64_t* stats, int32_t* nstg,
                     int64_t* work, int32_t* sens_iters, void* stream);

/* The same with parameter directions: Sp_out [M, n_par, Ny] = d y(tfinal) / d par[par_index[k]], k < n_par <= 4, for the
 * parameters of the handle's Hamiltonian (zero at t0).  The reference gets these by appending the parameter to the state
 * with zero dynamics and seeding its dual part (/root/reference/tests/cpp/ode_solver_test.cpp:630-640: y = [y1, y2, mu]);
 * here the source term d f / d par of the sensitivity equation comes from one more hyper-dual variable of H.  A control
 * law is held constant under differentiation (evalDynsU semantics).  n_par = 0: jb200_solve_sens.                     */
int jb200_solve_sens_par(void* handle, const double* y0, const double* tspan, const double* S0, int32_t n_par,
                         const int32_t* par_index, int64_t M, double* y_out, double* S_out, double* Sp_out, double* t_out,
                         double* h_out, int32_t* res, int64_t* stats, int32_t* nstg, int64_t* work, int32_t* sens_iters,
                         void* stream);

/* Explicit Dormand-Prince 8(5,3) step for non-stiff batches: replaces
 * Dopri853Te.step(htry) of /root/reference/src/python/dopri853te.py:290-315 (+ dy :207-287, error :191-205,
 * control :318-355) with derivs = the handle's Hamiltonian dynamics.  One call = one ACCEPTED step of every
 * trajectory (rejected attempts retried inside).  All pointers are device arrays updated in place: x [M],
 * y / dydx [M,Ny], errold [M] (initialise to 1e-4), hnext [M]; htry [M] in; hdid [M], nrej [M] (may be NULL) out.
 * norm_M = the batch size of the reference's error norm (0: M); aliased = 0 for the first step of a stepper
 * object, 1 afterwards (the reference's y / yout aliasing after its first step, :311); its float32
 * coefficients, non-growing hnext and beta = 0 are reproduced as well (csrc/dopri853.cuh).                     */
int jb200_dopri853_step(void* handle, int64_t M, int64_t norm_M, double atol, double rtol, int aliased, double* x, double* y,
                        double* dydx, double* errold, double* hnext, const double* htry, double* hdid, int32_t* nrej,
                        void* stream);

/* The same step with a dual part: replaces Dopri853TeD (/root/reference/src/cpp/dopri853ted.hpp:26-734, TensorDual state through
 * dy()).  yd [M,Ny,nd] (1 <= nd <= 8) is the dual part of y, updated in place across the accepted step: the stages are
 * differentiated one by one, K_s' = J(Y_s) (S + h sum_j a_sj K_j').  Step-size and error control look at the real part only:
 * x, y, dydx, hnext, hdid are those of jb200_dopri853_step bit for bit.                                                    */
int jb200_dopri853_step_dual(void* handle, int64_t M, int64_t norm_M, double atol, double rtol, int aliased, double* x, double* y,
                             double* dydx, double* errold, double* hnext, const double* htry, double* hdid, int32_t* nrej,
                             double* yd, int32_t nd, void* stream);

/* batched small dense LU factor + solve, n in {2,3,4} (one thread per system) or 16 (16 lanes per
 * system, one matrix row per lane): A [M,n,n], b [M,n] -> x [M,n].
 * Aim/bim/xim == NULL selects the real path, otherwise complex (separate re/im planes).    */
int jb200_lu_solve(int n, int64_t M, const double* Are, const double* Aim, const double* bre,
                   const double* bim, double* xre, double* xim, void* stream);

/* deterministic pow used by the controller (test hook): out[i] = x[i]^y[i], device pointers */
int jb200_det_pow(const double* x, const double* y, int64_t n, double* out, void* stream);

/* FP64 roofline probe: runs a dependent-chain-free DFMA kernel and returns achieved TFLOP/s */
int jb200_fp64_peak(int iters, double* tflops);
/* FP64 tensor-core (DMMA, mma.sync m8n8k4 f64) throughput probe in TFLOP/s, next to the DFMA probe: tcgen05 has no
 * f64 kind, so this is the only tensor-core rate an fp64 GEMM could reach on sm_100a.                         */
int jb200_dmma_peak(int iters, double* tflops);

const char* jb200_last_error(void);
const char* jb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* JANUS_B200_H */
