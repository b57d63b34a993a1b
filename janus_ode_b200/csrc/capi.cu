// C ABI (include/janus_b200.h) + the CUDA kernels of the Hamiltonian Radau path, sm_100a.
//
// Kernels:
//   radau_kernel<Ham,S,false>  fixed order: persistent fused integrator, one thread = one trajectory,
//                              lanes refill from a global work counter when their trajectory ends
//                              (active-set maintenance: finished trajectories stop costing work)
//   init/partition/epoch_kernel<Ham,S>  variable order: trajectories are regrouped by stage count
//                              between epochs of <= 12 attempted steps, index lists counting-sorted by a regime key, so every warp runs one
//                              attempt_step<S> instantiation (order-homogeneous warps); between
//                              epochs a trajectory is a 568-byte Parked record in HBM
//   dyns_kernel / jac_kernel   stage evaluation only (dual / hyper-dual), for the dyns /
//                              dyns.jacobian surface of the API and for parity tests
//   lu_kernel                  batched small dense LU factor+solve, real / complex
//   detpow_kernel, dfma_kernel test hook and FP64 roofline probe
// Compiled with -fmad=false: every fused multiply-add is an explicit fma() (arithmetic contract,
// see radau_core.cuh).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/janus_b200.h"
#include "radau_core.cuh"
#ifdef JB_USER_HAM_HEADER
// Plugin build (janus_ode_b200/user.py): the same library compiled around ONE user-supplied
// Hamiltonian struct `jb::UserHam` (interface: csrc/hamiltonians.cuh), exposed as JB200_HAM_USER.
#include JB_USER_HAM_HEADER
#endif

namespace jb {

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define JB_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(JB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
  } while (0)

// ------------------------------------------------------------------ kernels
template <int NY>
__device__ __forceinline__ void load_row(const double* __restrict__ src, long long m, double* dst) {
  if (NY % 2 == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(src + m * NY);
#pragma unroll
    for (int i = 0; i < NY / 2; i++) {
      const double2 v = __ldg(s2 + i);
      dst[2 * i] = v.x;
      dst[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NY; i++) dst[i] = __ldg(src + m * NY + i);
  }
}
template <int NY>
__device__ __forceinline__ void store_row(double* __restrict__ dst, long long m, const double* src) {
  if (NY % 2 == 0) {
    double2* d2 = reinterpret_cast<double2*>(dst + m * NY);
#pragma unroll
    for (int i = 0; i < NY / 2; i++) d2[i] = make_double2(src[2 * i], src[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < NY; i++) dst[m * NY + i] = src[i];
  }
}

#ifndef JB_MINB_FIXED
#define JB_MINB_FIXED 2   // resident CTAs/SM targeted for the fixed-order kernels (register budget)
#endif
#ifndef JB_MINB_VAR
#define JB_MINB_VAR 2
#endif
#ifndef JB_THREADS
#define JB_THREADS 128
#endif

// optional dense output (tspan with interior points): device arrays, nullptr when unused
struct DenseArgs {
  int ntspan;      // columns of tspan (>= 2)
  double* tout;    // [M, ntspan] or nullptr
  double* yout;    // [M, ntspan, NY]
  int* nout;       // [M]
};

template <class Ham, int SM, bool VAR>
__global__ void __launch_bounds__(JB_THREADS, VAR ? JB_MINB_VAR : JB_MINB_FIXED) radau_kernel(const __grid_constant__ RadauParams P,
                                                    const double* __restrict__ y0,
                                                    const double* __restrict__ tspan, long long M,
                                                    double* __restrict__ y_out, double* __restrict__ t_out,
                                                    double* __restrict__ h_out, int* __restrict__ res,
                                                    long long* __restrict__ stats, int* __restrict__ nstg,
                                                    long long* __restrict__ work, const DenseArgs dense,
                                                    unsigned long long* __restrict__ counter) {
  constexpr int NY = Ham::NV;
  Traj<NY, SM> T;
  long long m = -1;
  bool have = false;
  const int nts = dense.ntspan;
  for (;;) {
    if (!have) {
      m = (long long)atomicAdd(counter, 1ull);
      if (m >= M) break;
      double y[NY];
      load_row<NY>(y0, m, y);
      traj_init<Ham, NY, SM>(T, P, y, __ldg(tspan + m * nts), __ldg(tspan + m * nts + nts - 1));
      have = true;
    }
    const DenseOut dout{tspan + m * nts, nts, dense.tout ? dense.tout + m * nts : nullptr,
                        dense.yout ? dense.yout + m * nts * NY : nullptr};
    const DenseOut* dp = dense.tout ? &dout : nullptr;
    bool done;
    if (T.res == 0 && traj_active(T.PosNeg, T.t, T.tfinal, T.st[ST_STEP], P.MaxNbrStep)) {
      done = one_pass<Ham, NY, SM, VAR>(T, P, dp);
    } else {  // failed at init, or the step cap ended the loop (radau.m:1287)
      if (T.res == 0 && T.PosNeg * T.t < T.PosNeg * T.tfinal) T.res = 3;
      done = true;
    }
    if (done) {
      store_row<NY>(y_out, m, T.y);
      t_out[m] = T.t;
      h_out[m] = T.h;
      res[m] = T.res;
#pragma unroll
      for (int k = 0; k < 8; k++) stats[m * 8 + k] = (long long)T.st[k];
      if (nstg) nstg[m] = T.NbrStg;
      if (work) work[m] = T.work;
      if (dp) {
        dense_finish<NY, SM>(T, dout);
        dense.nout[m] = T.nout;
      }
      have = false;
    }
  }
}

// ---- variable-order path: order-homogeneous epochs over parked trajectories (radau_core.cuh) ----
// Regime key of a parked trajectory: stage count (2 bits) x step-size octave (6 bits) x Newton
// iterations of its last attempt (2 bits).  Trajectories with the same key behave alike (same number
// of Newton iterations, same accept/reject pattern), so warps built from key-sorted lists diverge
// less.  0 = finished.  Purely a scheduling hint: results do not depend on it.
constexpr int NKEYS = 4 * 256;
__device__ __forceinline__ unsigned short regime_key(int NbrStg, double h, int Newt) {
  const int k = (NbrStg - 1) >> 1;
  int e = (int)((__double_as_longlong(h) >> 52) & 0x7ff) - 1023 + 44;   // |h| in [2^-44, 2^20) -> 0..63
  e = e < 0 ? 0 : (e > 63 ? 63 : e);
  int nw = Newt - 1;
  nw = nw < 0 ? 0 : (nw > 3 ? 3 : nw);
  return (unsigned short)(1 + k * 256 + e * 4 + nw);
}

template <class Ham>
__global__ void init_kernel(const __grid_constant__ RadauParams P, const double* __restrict__ y0,
                            const double* __restrict__ tspan, int nts, long long M,
                            Parked<Ham::NV>* __restrict__ st, unsigned short* __restrict__ tag) {
  constexpr int NY = Ham::NV;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    double y[NY];
    load_row<NY>(y0, m, y);
    const double t0 = __ldg(tspan + m * nts);
    Traj<NY, 7> T;
    traj_init<Ham, NY, 7>(T, P, y, t0, __ldg(tspan + m * nts + nts - 1));
    park<NY, 7>(T, t0, false, false, st[m]);
    tag[m] = regime_key(T.NbrStg, T.h, T.Newt);
  }
}

__global__ void hist_kernel(const unsigned short* __restrict__ tag, long long M, int* __restrict__ counts) {
  __shared__ int sh[NKEYS];
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const int t = tag[m];
    if (t) atomicAdd(&sh[t - 1], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x)
    if (sh[i]) atomicAdd(&counts[i], sh[i]);
}

// counting-sort scatter: cursor[key] starts at the exclusive prefix sum of the histogram.
// Warp-aggregated: lanes holding the same key are matched and claim their slots with ONE atomic
// (the live keys are few, so per-lane atomics serialise on a handful of addresses).
__global__ void scatter_kernel(const unsigned short* __restrict__ tag, long long M, int* __restrict__ cursor,
                               int* __restrict__ list) {
  const int lane = threadIdx.x & 31;
  for (long long base = blockIdx.x * (long long)blockDim.x; base < M; base += (long long)gridDim.x * blockDim.x) {
    const long long m = base + threadIdx.x;
    const int t = (m < M) ? (int)tag[m] : 0;
    const unsigned peers = __match_any_sync(0xffffffffu, t);
    if (t) {
      const int leader = __ffs(peers) - 1;
      int pos = 0;
      if (lane == leader) pos = atomicAdd(&cursor[t - 1], __popc(peers));
      pos = __shfl_sync(peers, pos, leader);
      list[pos + __popc(peers & ((1u << lane) - 1u))] = (int)m;
    }
  }
}

#ifndef JB_MINB_HI
#define JB_MINB_HI 2      // resident CTAs/SM for the 5- and 7-stage epoch kernels (state spills to local memory)
#endif
#ifndef JB_THREADS_HI
#define JB_THREADS_HI 128
#endif
template <int S>
struct EpochCfg {
  static constexpr int threads = S >= 5 ? JB_THREADS_HI : JB_THREADS;
  static constexpr int minb = S >= 5 ? JB_MINB_HI : JB_MINB_FIXED;
};

template <class Ham, int S, bool VAR>
__global__ void __launch_bounds__(EpochCfg<S>::threads, EpochCfg<S>::minb)
    epoch_kernel(const __grid_constant__ RadauParams P, Parked<Ham::NV>* __restrict__ st, const int* __restrict__ list,
                 int count, int max_passes, unsigned long long* __restrict__ counter, unsigned short* __restrict__ tag,
                 double* __restrict__ y_out, double* __restrict__ t_out, double* __restrict__ h_out,
                 int* __restrict__ res, long long* __restrict__ stats, int* __restrict__ nstg,
                 long long* __restrict__ work, const double* __restrict__ tspan, const DenseArgs dense) {
  constexpr int NY = Ham::NV;
  const int nts = dense.ntspan;
  for (;;) {
    const long long i = (long long)atomicAdd(counter, 1ull);
    if (i >= count) break;
    const long long m = list[i];
    Traj<NY, S> T;
    double t0;
    bool head_done;
    unpark<NY, S>(st[m], T, t0, head_done);
    const DenseOut dout{tspan + m * nts, nts, dense.tout ? dense.tout + m * nts : nullptr,
                        dense.yout ? dense.yout + m * nts * NY : nullptr};
    const DenseOut* dp = dense.tout ? &dout : nullptr;
    const bool fin = run_epoch<Ham, NY, S, VAR>(T, P, head_done, max_passes, dp);
    park<NY, S>(T, t0, head_done, fin, st[m]);
    tag[m] = fin ? (unsigned short)0 : regime_key(T.NbrStg, T.h, T.Newt);
    if (fin) {
      store_row<NY>(y_out, m, T.y);
      t_out[m] = T.t;
      h_out[m] = T.h;
      res[m] = T.res;
#pragma unroll
      for (int k = 0; k < 8; k++) stats[m * 8 + k] = (long long)T.st[k];
      if (nstg) nstg[m] = T.NbrStg;
      if (work) work[m] = T.work;
      if (dp) {
        dense_finish<NY, S>(T, dout);
        dense.nout[m] = T.nout;
      }
    }
  }
}

template <class Ham>
__global__ void dyns_kernel(const __grid_constant__ RadauParams P, const double* __restrict__ y, long long M,
                            double* __restrict__ f) {
  constexpr int NY = Ham::NV;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    double yy[NY], ff[NY];
    load_row<NY>(y, m, yy);
    ham_dyns<Ham>(yy, P.ham_par, P.sign, ff);
    store_row<NY>(f, m, ff);
  }
}

template <class Ham>
__global__ void jac_kernel(const __grid_constant__ RadauParams P, const double* __restrict__ y, long long M,
                           double* __restrict__ J) {
  constexpr int NY = Ham::NV;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    double yy[NY], JJ[NY * NY];
    load_row<NY>(y, m, yy);
    ham_jac<Ham>(yy, P.ham_par, P.sign, JJ);
#pragma unroll
    for (int r = 0; r < NY; r++) store_row<NY>(J, m * NY + r, &JJ[r * NY]);
  }
}

template <int N, bool CPLX>
__global__ void lu_kernel(long long M, const double* __restrict__ Are, const double* __restrict__ Aim,
                          const double* __restrict__ bre, const double* __restrict__ bim, double* __restrict__ xre,
                          double* __restrict__ xim) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    if (!CPLX) {
      RLU<N> L;
#pragma unroll
      for (int k = 0; k < N * N; k++) L.a[k] = Are[m * N * N + k];
      rlu_factor<N>(L);
      double b[N];
#pragma unroll
      for (int k = 0; k < N; k++) b[k] = bre[m * N + k];
      rlu_solve<N>(L, b);
#pragma unroll
      for (int k = 0; k < N; k++) xre[m * N + k] = b[k];
    } else {
      ZLU<N> L;
#pragma unroll
      for (int k = 0; k < N * N; k++) {
        L.re[k] = Are[m * N * N + k];
        L.im[k] = Aim[m * N * N + k];
      }
      zlu_factor<N>(L);
      double br[N], bi[N];
#pragma unroll
      for (int k = 0; k < N; k++) {
        br[k] = bre[m * N + k];
        bi[k] = bim[m * N + k];
      }
      zlu_solve<N>(L, br, bi);
#pragma unroll
      for (int k = 0; k < N; k++) {
        xre[m * N + k] = br[k];
        xim[m * N + k] = bi[k];
      }
    }
  }
}

__global__ void detpow_kernel(const double* x, const double* y, long long n, double* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = det_pow(x[i], y[i]);
}

// FP64 roofline probe: 16 independent DFMA chains per thread
__global__ void dfma_kernel(int iters, double seed, double* sink) {
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = seed + k * 1e-3 + threadIdx.x * 1e-6;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = fma(a[k], b, c);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  if (s == 123.456) sink[0] = s;
}

// ------------------------------------------------------------------ host side
struct Handle {
  int ham_id = 0, ny = 0, device = 0, sm_count = 148;
  RadauParams P;
  int kernel_sm = 3;     // SM template value in use
  bool kernel_var = false;
  unsigned long long* d_counter = nullptr;   // [4] work-fetch counters (one per stage count)
  int launches = 0;
  // variable-order epoch workspace
  long long vcap = 0;
  void* d_parked = nullptr;
  unsigned short* d_tag = nullptr;
  int* d_lists = nullptr;
  int* d_counts = nullptr;
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // the per-order epoch kernels of one epoch overlap
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  bool fixed_epochs = false;
  int epoch_passes = 12;   // attempted steps per trajectory per epoch (flat optimum 8..24 on B200)
  // workspace of jb200_solve_host
  long long cap = 0;
  double *d_y0 = nullptr, *d_ts = nullptr, *d_y = nullptr, *d_t = nullptr, *d_h = nullptr;
  int *d_res = nullptr, *d_nstg = nullptr;
  long long* d_stats = nullptr;
  long long* d_work = nullptr;
  cudaStream_t own_stream = nullptr;
};

static int ham_ny(int ham_id) {
#ifdef JB_USER_HAM_HEADER
  if (ham_id == JB200_HAM_USER) return UserHam::NV;
#endif
  switch (ham_id) {
    case JB200_HAM_HARMONIC: return 2;
    case JB200_HAM_VDP:
    case JB200_HAM_VDP_PMP_SMOOTH:
    case JB200_HAM_VDP_PMP_BANG:
    case JB200_HAM_VDP_AUG_BANG: return 4;
    default: return -1;
  }
}
static int ham_npar(int ham_id) {
#ifdef JB_USER_HAM_HEADER
  if (ham_id == JB200_HAM_USER) return UserHam::NPAR;
#endif
  switch (ham_id) {
    case JB200_HAM_HARMONIC: return 1;
    case JB200_HAM_VDP: return 1;
    case JB200_HAM_VDP_PMP_SMOOTH: return 1;
    case JB200_HAM_VDP_PMP_BANG: return 3;
    case JB200_HAM_VDP_AUG_BANG: return 2;
    default: return -1;
  }
}

// option checks of the reference constructor (radaute_impl.hpp:65-214), returned not exit()ed
static int validate(const jb200_opts* o, int ny) {
  if (o->n_tol != 1 && o->n_tol != ny) return fail(JB200_EINVAL, "RelTol/AbsTol: vector of length 1 or Ny");
  for (int i = 0; i < o->n_tol; i++) {
    if (!(o->AbsTol[i] >= 0)) return fail(JB200_EINVAL, "Wrong input AbsTol must be a positive number");
    if (!(o->RelTol[i] >= 10 * EPS)) return fail(JB200_EINVAL, "Relative tolerance are too small.");
  }
  if (!(o->JacRecompute >= 0)) return fail(JB200_EINVAL, "JacRecompute must be positive");
  if (o->JacRecompute >= 1.0) return fail(JB200_EINVAL, "Invalid JacRecompute must be less than one");
  if (o->MaxNbrNewton < 4) return fail(JB200_EINVAL, "MaxNbrNewton integer >= 4");
  const int stg[3] = {o->NbrStg, o->MinNbrStg, o->MaxNbrStg};
  const char* nm[3] = {"NbrStg", "MinNbrStg", "MaxNbrStg"};
  for (int k = 0; k < 3; k++)
    if (stg[k] != 1 && stg[k] != 3 && stg[k] != 5 && stg[k] != 7)
      return fail(JB200_EINVAL, std::string(nm[k]) + " must be 1, 3, 5 or 7");
  if (o->NbrStg < o->MinNbrStg || o->NbrStg > o->MaxNbrStg)
    return fail(JB200_EINVAL, "NbrStg must be between MinNbrStg and MaxNbrStg");
  if (!(o->Safe > 0.001 && o->Safe < 1)) return fail(JB200_EINVAL, "Curious input for Safe, must be in ]0.001 .. 1[");
  if (o->Quot1 > 1) return fail(JB200_EINVAL, "Curious input for Quot1, must be <=1");
  if (o->Quot2 < 1) return fail(JB200_EINVAL, "Curious input for Quot2, must be >1");
  if (!(o->FacL <= 1.0) || !(o->FacL > 0)) return fail(JB200_EINVAL, "Curious input for FacL default is 0.2");
  if (!(o->FacR >= 1.0)) return fail(JB200_EINVAL, "Curious input for FacR default is 8.0");
  if (!(o->Vitu >= 0)) return fail(JB200_EINVAL, "Curious input for Vitu default is 0.002");
  if (!(o->Vitd >= 0)) return fail(JB200_EINVAL, "Curious input for Vitd default is 0.8");
  if (!(o->hhou >= 1)) return fail(JB200_EINVAL, "Curious input for hhou default is 1.2");
  if (!(o->hhod >= 0)) return fail(JB200_EINVAL, "Curious input for hhod default is 0.8");
  if (o->costate_sign != 1 && o->costate_sign != -1) return fail(JB200_EINVAL, "costate_sign must be +1 or -1");
  if (o->MaxNbrStep < 1 || o->MaxNbrStep > 2000000000LL) return fail(JB200_EINVAL, "MaxNbrStep out of range");
  return 0;
}

// per-stage-count constants (radau.m:858-862, 897-900), computed once on the host
static void fill_params(RadauParams& P, const jb200_opts* o, int ny, const double* hp, int nhp) {
  std::memset(&P, 0, sizeof P);
  for (int k = 0; k < 4; k++) {
    const int s = 2 * k + 1;
    const double ExpmNs = (s + 1.0) / (2.0 * s);
    double mrt = INFINITY;
    for (int i = 0; i < ny; i++) {
      const double rt = o->RelTol[o->n_tol == 1 ? 0 : i], at = o->AbsTol[o->n_tol == 1 ? 0 : i];
      const double quot = at / rt;
      P.RelTol1[k][i] = 0.1 * std::pow(rt, ExpmNs);
      P.AbsTol1[k][i] = P.RelTol1[k][i] * quot;
      mrt = std::fmin(mrt, P.RelTol1[k][i]);
    }
    P.FNewt[k] = std::fmax(10 * EPS / mrt, std::fmin(0.03, std::pow(mrt, 1.0 / ExpmNs - 1.0)));
    if (s == 1) P.FNewt[k] = std::fmax(10 * EPS / mrt, 0.03);
    P.inv_sqrt_sn[k] = 1.0 / std::sqrt((double)(s * ny));
  }
  P.inv_sqrt_n = 1.0 / std::sqrt((double)ny);
  P.InitialStep = o->InitialStep;
  P.Thet = o->JacRecompute;
  P.Safe = o->Safe;
  P.Quot1 = o->Quot1;
  P.Quot2 = o->Quot2;
  P.FacL = 1.0 / o->FacL;  // radaute_impl.hpp:185,191
  P.FacR = 1.0 / o->FacR;
  P.Vitu = o->Vitu;
  P.Vitd = o->Vitd;
  P.hhou = o->hhou;
  P.hhod = o->hhod;
  P.sign = (double)o->costate_sign;
  P.MaxNbrStep = o->MaxNbrStep;
  P.Start_Newt = o->Start_Newt;
  P.MaxNbrNewton = o->MaxNbrNewton;
  P.NbrStg = o->NbrStg;
  P.MinNbrStg = o->MinNbrStg;
  P.MaxNbrStg = o->MaxNbrStg;
  P.Gustafsson = o->Gustafsson;
  for (int i = 0; i < nhp && i < JB200_MAX_HAM_PARAMS; i++) P.ham_par[i] = hp[i];
}

template <class Ham, int SM, bool VAR>
static int launch_radau(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                        double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense, cudaStream_t st) {
  auto kern = radau_kernel<Ham, SM, VAR>;
  int bps = 0;
  JB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, JB_THREADS, 0));
  if (bps < 1) bps = 1;
  long long blocks = (long long)H->sm_count * bps;
  const long long need = (M + JB_THREADS - 1) / JB_THREADS;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  JB_CUDA(cudaMemsetAsync(H->d_counter, 0, sizeof(unsigned long long), st));
  kern<<<(unsigned)blocks, JB_THREADS, 0, st>>>(H->P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, H->d_counter);
  JB_CUDA(cudaGetLastError());
  H->launches = 1;
  return 0;
}

template <class Ham, int S, bool VAR>
static int launch_epoch(Handle* H, int k, int base, int count, long long M, double* y_out, double* t_out, double* h_out, int* res,
                        long long* stats, int* nstg, long long* work, const double* tspan, const DenseArgs& dense, cudaStream_t main_st, int& used) {
  // first non-empty bucket runs on the caller's stream, the others fork onto auxiliary streams
  cudaStream_t st = main_st;
  if (used > 0) {
    st = H->aux[used - 1];
    JB_CUDA(cudaStreamWaitEvent(st, H->ev_fork, 0));
  }
  auto kern = epoch_kernel<Ham, S, VAR>;
  constexpr int TH = EpochCfg<S>::threads;
  int bps = 0;
  JB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TH, 0));
  if (bps < 1) bps = 1;
  if (S >= 5 && bps > EpochCfg<S>::minb) bps = EpochCfg<S>::minb;
  long long blocks = (long long)H->sm_count * bps;
  const long long need = ((long long)count + TH - 1) / TH;
  if (blocks > need) blocks = need;
  kern<<<(unsigned)blocks, TH, 0, st>>>(H->P, (Parked<Ham::NV>*)H->d_parked, H->d_lists + base, count,
                                               H->epoch_passes, H->d_counter + k, H->d_tag, y_out, t_out, h_out, res,
                                               stats, nstg, work, tspan, dense);
  JB_CUDA(cudaGetLastError());
  H->launches += 1;
  if (used > 0) {
    JB_CUDA(cudaEventRecord(H->ev_join[used - 1], st));
    JB_CUDA(cudaStreamWaitEvent(main_st, H->ev_join[used - 1], 0));
  }
  used += 1;
  return 0;
}

// variable order: init -> { partition by stage count ; one order-homogeneous epoch kernel per
// non-empty bucket } until no trajectory is left.  Synchronises the stream once per epoch (4 ints).
template <class Ham, bool VAR = true>
static int solve_var(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                     double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense, cudaStream_t st) {
  constexpr int NY = Ham::NV;
  if (M > 2000000000LL) return fail(JB200_EINVAL, "M too large for the variable-order path");
  if (M > H->vcap) {
    cudaFree(H->d_parked); cudaFree(H->d_tag); cudaFree(H->d_lists); cudaFree(H->d_counts);
    H->vcap = 0;
    JB_CUDA(cudaMalloc(&H->d_parked, sizeof(Parked<NY>) * M));
    JB_CUDA(cudaMalloc(&H->d_tag, sizeof(unsigned short) * M));
    JB_CUDA(cudaMalloc(&H->d_lists, sizeof(int) * M));
    JB_CUDA(cudaMalloc(&H->d_counts, sizeof(int) * NKEYS));
    H->vcap = M;
  }
  const long long gblocks = std::max<long long>(1, std::min<long long>((M + 255) / 256, (long long)H->sm_count * 8));
  init_kernel<Ham><<<(unsigned)gblocks, 256, 0, st>>>(H->P, y0, tspan, dense.ntspan, M, (Parked<NY>*)H->d_parked, H->d_tag);
  JB_CUDA(cudaGetLastError());
  H->launches = 1;
  std::vector<int> hist(NKEYS);
  for (long long epoch = 0;; epoch++) {
    // counting sort of the live trajectories by regime key -> one contiguous index list, order-major
    JB_CUDA(cudaMemsetAsync(H->d_counts, 0, sizeof(int) * NKEYS, st));
    JB_CUDA(cudaMemsetAsync(H->d_counter, 0, sizeof(unsigned long long) * 4, st));
    hist_kernel<<<(unsigned)gblocks, 256, 0, st>>>(H->d_tag, M, H->d_counts);
    JB_CUDA(cudaGetLastError());
    JB_CUDA(cudaMemcpyAsync(hist.data(), H->d_counts, sizeof(int) * NKEYS, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    int cnt[4] = {0, 0, 0, 0}, base[4] = {0, 0, 0, 0}, run = 0;
    for (int i = 0; i < NKEYS; i++) {
      const int c = hist[i];
      hist[i] = run;           // exclusive prefix sum = scatter cursor
      if ((i & 255) == 0) base[i >> 8] = run;
      cnt[i >> 8] += c;
      run += c;
    }
    if (run == 0) break;
    JB_CUDA(cudaMemcpyAsync(H->d_counts, hist.data(), sizeof(int) * NKEYS, cudaMemcpyHostToDevice, st));
    scatter_kernel<<<(unsigned)gblocks, 256, 0, st>>>(H->d_tag, M, H->d_counts, H->d_lists);
    JB_CUDA(cudaGetLastError());
    H->launches += 2;
    int rc = 0, used = 0;
    JB_CUDA(cudaEventRecord(H->ev_fork, st));
    // largest stage count first: its kernel is the longest-running per trajectory
    if (cnt[3]) rc = launch_epoch<Ham, 7, VAR>(H, 3, base[3], cnt[3], M, y_out, t_out, h_out, res, stats, nstg, work, tspan, dense, st, used);
    if (!rc && cnt[2]) rc = launch_epoch<Ham, 5, VAR>(H, 2, base[2], cnt[2], M, y_out, t_out, h_out, res, stats, nstg, work, tspan, dense, st, used);
    if (!rc && cnt[1]) rc = launch_epoch<Ham, 3, VAR>(H, 1, base[1], cnt[1], M, y_out, t_out, h_out, res, stats, nstg, work, tspan, dense, st, used);
    if (!rc && cnt[0]) rc = launch_epoch<Ham, 1, VAR>(H, 0, base[0], cnt[0], M, y_out, t_out, h_out, res, stats, nstg, work, tspan, dense, st, used);
    if (rc) return rc;
  }
  return 0;
}

template <class Ham>
static int dispatch_order(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                          double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense, cudaStream_t st) {
  const RadauParams& P = H->P;
  if (P.MinNbrStg == P.MaxNbrStg && H->fixed_epochs)   // fixed order through the regime-sorted epochs
    return solve_var<Ham, false>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
  if (P.MinNbrStg == P.MaxNbrStg && P.MaxNbrStg == 3)
    return launch_radau<Ham, 3, false>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
  if (P.MinNbrStg == P.MaxNbrStg && P.MaxNbrStg == 1)
    return launch_radau<Ham, 1, false>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
  if (P.MinNbrStg == P.MaxNbrStg && P.MaxNbrStg == 5)
    return launch_radau<Ham, 5, false>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
  if (P.MinNbrStg == P.MaxNbrStg && P.MaxNbrStg == 7)
    return launch_radau<Ham, 7, false>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
  return solve_var<Ham>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
}

#if defined(JB_USER_HAM_HEADER)
#define JB_HAM_SWITCH(ID, CALL)                                              \
  switch (ID) {                                                              \
    case JB200_HAM_USER: { using HamT = UserHam; return CALL; }              \
    default: return fail(JB200_ENOTSUP, "plugin library: JB200_HAM_USER only"); \
  }
#elif defined(JB_FAST_BUILD)   /* developer builds: van der Pol only (5x faster to compile) */
#define JB_HAM_SWITCH(ID, CALL)                                              \
  switch (ID) {                                                              \
    case JB200_HAM_VDP: { using HamT = HamVdp; return CALL; }                \
    default: return fail(JB200_ENOTSUP, "fast build: van der Pol only");     \
  }
#else
#define JB_HAM_SWITCH(ID, CALL)                                              \
  switch (ID) {                                                              \
    case JB200_HAM_HARMONIC: { using HamT = HamHarmonic; return CALL; }      \
    case JB200_HAM_VDP: { using HamT = HamVdp; return CALL; }                \
    case JB200_HAM_VDP_PMP_SMOOTH: { using HamT = HamVdpPmp<false>; return CALL; } \
    case JB200_HAM_VDP_PMP_BANG: { using HamT = HamVdpPmp<true>; return CALL; }    \
    case JB200_HAM_VDP_AUG_BANG: { using HamT = HamVdpAugBang; return CALL; }      \
    default: return fail(JB200_ENOTSUP, "unknown Hamiltonian id");           \
  }
#endif

static int solve_dev(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                     double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense, cudaStream_t st) {
  if (M <= 0) {
    H->launches = 0;
    return 0;
  }
  JB_HAM_SWITCH(H->ham_id, (dispatch_order<HamT>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st)));
}

template <class Ham>
static int launch_dyns(Handle* H, const double* y, long long M, double* f, cudaStream_t st) {
  const long long blocks = std::min<long long>((M + 127) / 128, (long long)H->sm_count * 16);
  dyns_kernel<Ham><<<(unsigned)std::max<long long>(blocks, 1), 128, 0, st>>>(H->P, y, M, f);
  JB_CUDA(cudaGetLastError());
  return 0;
}
template <class Ham>
static int launch_jac(Handle* H, const double* y, long long M, double* J, cudaStream_t st) {
  const long long blocks = std::min<long long>((M + 127) / 128, (long long)H->sm_count * 16);
  jac_kernel<Ham><<<(unsigned)std::max<long long>(blocks, 1), 128, 0, st>>>(H->P, y, M, J);
  JB_CUDA(cudaGetLastError());
  return 0;
}

template <int N>
static int launch_lu(long long M, const double* Are, const double* Aim, const double* bre, const double* bim,
                     double* xre, double* xim, cudaStream_t st) {
  const long long blocks = std::max<long long>(1, std::min<long long>((M + 127) / 128, 148 * 16));
  if (Aim)
    lu_kernel<N, true><<<(unsigned)blocks, 128, 0, st>>>(M, Are, Aim, bre, bim, xre, xim);
  else
    lu_kernel<N, false><<<(unsigned)blocks, 128, 0, st>>>(M, Are, nullptr, bre, nullptr, xre, nullptr);
  JB_CUDA(cudaGetLastError());
  return 0;
}

static int ensure_workspace(Handle* H, long long M) {
  if (M <= H->cap) return 0;
  cudaFree(H->d_y0); cudaFree(H->d_ts); cudaFree(H->d_y); cudaFree(H->d_t); cudaFree(H->d_h);
  cudaFree(H->d_res); cudaFree(H->d_nstg); cudaFree(H->d_stats); cudaFree(H->d_work);
  H->cap = 0;
  const int ny = H->ny;
  JB_CUDA(cudaMalloc(&H->d_y0, sizeof(double) * M * ny));
  JB_CUDA(cudaMalloc(&H->d_ts, sizeof(double) * M * 2));
  JB_CUDA(cudaMalloc(&H->d_y, sizeof(double) * M * ny));
  JB_CUDA(cudaMalloc(&H->d_t, sizeof(double) * M));
  JB_CUDA(cudaMalloc(&H->d_h, sizeof(double) * M));
  JB_CUDA(cudaMalloc(&H->d_res, sizeof(int) * M));
  JB_CUDA(cudaMalloc(&H->d_nstg, sizeof(int) * M));
  JB_CUDA(cudaMalloc(&H->d_stats, sizeof(long long) * M * 8));
  JB_CUDA(cudaMalloc(&H->d_work, sizeof(long long) * M));
  H->cap = M;
  return 0;
}

}  // namespace jb

using namespace jb;

extern "C" {

int jb200_opts_default(jb200_opts* o) {
  if (!o) return fail(JB200_EINVAL, "null opts");
  std::memset(o, 0, sizeof *o);
  o->RelTol[0] = 1.0e-3;
  o->AbsTol[0] = 1.0e-6;
  o->n_tol = 1;
  o->InitialStep = 1.0e-2;
  o->MaxStep = 1.0;
  o->JacRecompute = 1.0e-3;
  o->Start_Newt = 0;
  o->MaxNbrNewton = 7;
  o->NbrStg = 3;
  o->MinNbrStg = 3;
  o->MaxNbrStg = 7;
  o->MaxNbrStep = 10000;
  o->Safe = 0.9;
  o->Quot1 = 1.0;
  o->Quot2 = 1.2;
  o->FacL = 0.2;
  o->FacR = 8.0;
  o->Vitu = 0.002;
  o->Vitd = 0.8;
  o->hhou = 1.2;
  o->hhod = 0.8;
  o->Gustafsson = 1;
  o->costate_sign = 1;
  return 0;
}

int jb200_create(const jb200_opts* o, int ham_id, const double* ham_params, int n_ham_params, void** handle) {
  if (!o || !handle) return fail(JB200_EINVAL, "null argument");
  const int ny = ham_ny(ham_id);
  if (ny < 0) return fail(JB200_ENOTSUP, "unknown Hamiltonian id");
  if (n_ham_params < ham_npar(ham_id) || (n_ham_params > 0 && !ham_params))
    return fail(JB200_EINVAL, "Hamiltonian needs " + std::to_string(ham_npar(ham_id)) + " parameters");
  int rc = validate(o, ny);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(JB200_ECUDA, "no CUDA device: janus_b200 has no CPU fallback");
  Handle* H = new Handle();
  H->ham_id = ham_id;
  H->ny = ny;
  fill_params(H->P, o, ny, ham_params, n_ham_params);
  cudaError_t e = cudaGetDevice(&H->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&H->sm_count, cudaDevAttrMultiProcessorCount, H->device);
  if (e == cudaSuccess) e = cudaMalloc(&H->d_counter, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&H->own_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 3 && e == cudaSuccess; i++) {
    e = cudaStreamCreateWithFlags(&H->aux[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&H->ev_join[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&H->ev_fork, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete H;
    return fail(JB200_ECUDA, std::string("jb200_create: ") + cudaGetErrorString(e));
  }
  if (const char* ep = std::getenv("JB200_EPOCH_PASSES")) H->epoch_passes = std::max(1, atoi(ep));
  if (const char* fe = std::getenv("JB200_FIXED_EPOCHS")) H->fixed_epochs = atoi(fe) != 0;
  *handle = H;
  return 0;
}

void jb200_destroy(void* handle) {
  Handle* H = (Handle*)handle;
  if (!H) return;
  cudaFree(H->d_counter);
  cudaFree(H->d_parked); cudaFree(H->d_tag); cudaFree(H->d_lists); cudaFree(H->d_counts);
  cudaFree(H->d_y0); cudaFree(H->d_ts); cudaFree(H->d_y); cudaFree(H->d_t); cudaFree(H->d_h);
  cudaFree(H->d_res); cudaFree(H->d_nstg); cudaFree(H->d_stats); cudaFree(H->d_work);
  if (H->own_stream) cudaStreamDestroy(H->own_stream);
  for (int i = 0; i < 3; i++) {
    if (H->aux[i]) cudaStreamDestroy(H->aux[i]);
    if (H->ev_join[i]) cudaEventDestroy(H->ev_join[i]);
  }
  if (H->ev_fork) cudaEventDestroy(H->ev_fork);
  delete H;
}

int jb200_ny(void* handle) { return handle ? ((Handle*)handle)->ny : JB200_EINVAL; }

int jb200_solve(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out, double* t_out,
                double* h_out, int32_t* res, int64_t* stats, int32_t* nstg, int64_t* work, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M < 0) return fail(JB200_EINVAL, "M < 0");
  if (M > 0 && (!y0 || !tspan || !y_out || !t_out || !h_out || !res || !stats)) return fail(JB200_EINVAL, "null buffer");
  const DenseArgs none{2, nullptr, nullptr, nullptr};
  return solve_dev(H, y0, tspan, (long long)M, y_out, t_out, h_out, (int*)res, (long long*)stats, (int*)nstg,
                   (long long*)work, none, (cudaStream_t)stream);
}

int jb200_solve_dense(void* handle, const double* y0, const double* tspan, int ntspan, int64_t M, double* y_out,
                      double* t_out, double* h_out, int32_t* res, int64_t* stats, int32_t* nstg, int64_t* work,
                      double* tout, double* yout, int32_t* nout, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M < 0) return fail(JB200_EINVAL, "M < 0");
  if (ntspan < 2) return fail(JB200_EINVAL, "tspan needs at least [t0, tfinal]");
  if (M > 0 && (!y0 || !tspan || !y_out || !t_out || !h_out || !res || !stats || !tout || !yout || !nout))
    return fail(JB200_EINVAL, "null buffer");
  const DenseArgs dense{ntspan, tout, yout, (int*)nout};
  return solve_dev(H, y0, tspan, (long long)M, y_out, t_out, h_out, (int*)res, (long long*)stats, (int*)nstg,
                   (long long*)work, dense, (cudaStream_t)stream);
}

int jb200_solve_host(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out, double* t_out,
                     double* h_out, int32_t* res, int64_t* stats, int32_t* nstg, int64_t* work) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M < 0) return fail(JB200_EINVAL, "M < 0");
  if (M == 0) {
    H->launches = 0;
    return 0;
  }
  if (!y0 || !tspan || !y_out || !t_out || !h_out || !res || !stats) return fail(JB200_EINVAL, "null buffer");
  // tspan must be strictly monotonic (radaute_impl.hpp:54-59) -- checked here on the host copy
  for (int64_t m = 0; m < M; m++)
    if (!(tspan[2 * m + 1] != tspan[2 * m])) return fail(JB200_EINVAL, "Time vector must be strictly monotonic");
  int rc = ensure_workspace(H, M);
  if (rc) return rc;
  cudaStream_t st = H->own_stream;
  const int ny = H->ny;
  JB_CUDA(cudaMemcpyAsync(H->d_y0, y0, sizeof(double) * M * ny, cudaMemcpyHostToDevice, st));
  JB_CUDA(cudaMemcpyAsync(H->d_ts, tspan, sizeof(double) * M * 2, cudaMemcpyHostToDevice, st));
  const DenseArgs none{2, nullptr, nullptr, nullptr};
  rc = solve_dev(H, H->d_y0, H->d_ts, M, H->d_y, H->d_t, H->d_h, H->d_res, H->d_stats, H->d_nstg, H->d_work, none, st);
  if (rc) return rc;
  JB_CUDA(cudaMemcpyAsync(y_out, H->d_y, sizeof(double) * M * ny, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaMemcpyAsync(t_out, H->d_t, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaMemcpyAsync(h_out, H->d_h, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaMemcpyAsync(res, H->d_res, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaMemcpyAsync(stats, H->d_stats, sizeof(long long) * M * 8, cudaMemcpyDeviceToHost, st));
  if (nstg) JB_CUDA(cudaMemcpyAsync(nstg, H->d_nstg, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
  if (work) JB_CUDA(cudaMemcpyAsync(work, H->d_work, sizeof(long long) * M, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int jb200_last_launches(void* handle) { return handle ? ((Handle*)handle)->launches : 0; }

int jb200_eval_dyns(void* handle, const double* y, int64_t M, double* ydot, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  JB_HAM_SWITCH(H->ham_id, (launch_dyns<HamT>(H, y, M, ydot, (cudaStream_t)stream)));
}

int jb200_eval_jac(void* handle, const double* y, int64_t M, double* jac, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  JB_HAM_SWITCH(H->ham_id, (launch_jac<HamT>(H, y, M, jac, (cudaStream_t)stream)));
}

int jb200_lu_solve(int n, int64_t M, const double* Are, const double* Aim, const double* bre, const double* bim,
                   double* xre, double* xim, void* stream) {
  if (M <= 0) return 0;
  if (!Are || !bre || !xre) return fail(JB200_EINVAL, "null buffer");
  if (Aim && (!bim || !xim)) return fail(JB200_EINVAL, "complex path needs bim and xim");
  cudaStream_t st = (cudaStream_t)stream;
  switch (n) {
    case 2: return launch_lu<2>(M, Are, Aim, bre, bim, xre, xim, st);
    case 3: return launch_lu<3>(M, Are, Aim, bre, bim, xre, xim, st);
    case 4: return launch_lu<4>(M, Are, Aim, bre, bim, xre, xim, st);
    default: return fail(JB200_ENOTSUP, "jb200_lu_solve: n must be 2, 3 or 4");
  }
}

int jb200_det_pow(const double* x, const double* y, int64_t n, double* out, void* stream) {
  if (n <= 0) return 0;
  detpow_kernel<<<(unsigned)std::min<long long>((n + 127) / 128, 148 * 8), 128, 0, (cudaStream_t)stream>>>(x, y, n, out);
  JB_CUDA(cudaGetLastError());
  return 0;
}

int jb200_fp64_peak(int iters, double* tflops) {
  if (!tflops || iters <= 0) return fail(JB200_EINVAL, "bad argument");
  int dev = 0, sms = 148;
  JB_CUDA(cudaGetDevice(&dev));
  JB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  JB_CUDA(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  JB_CUDA(cudaEventCreate(&e0));
  JB_CUDA(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  dfma_kernel<<<blocks, threads>>>(iters, 1.0, sink);  // warm-up (clock ramp)
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {  // burst figure: best of 5
    JB_CUDA(cudaEventRecord(e0));
    dfma_kernel<<<blocks, threads>>>(iters, 1.0, sink);
    JB_CUDA(cudaEventRecord(e1));
    JB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    JB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  *tflops = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

const char* jb200_last_error(void) { return g_err.c_str(); }
const char* jb200_version(void) { return "janus_b200 0.1 (sm_100a)"; }

}  // extern "C"
