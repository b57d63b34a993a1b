// Kernels and host launchers of the Hamiltonian Radau path, templated on the Hamiltonian.
// Included by one translation unit per Hamiltonian (tu_ham.cu, compiled in parallel by
// __graft_entry__.build()) and by capi.cu in monolithic builds (plugins around a user Hamiltonian).
//
// Kernels:
//   radau_kernel<Ham,S,false,DENSE>  fixed order: persistent fused integrator, one thread = one
//                              trajectory, lanes refill from a global work counter when their
//                              trajectory ends (finished trajectories stop costing work)
//   init / hist / scatter / epoch_kernel<Ham,S,VAR,DENSE>  variable order: trajectories are regrouped
//                              by stage count between epochs of <= 12 attempted steps, index lists
//                              counting-sorted by a regime key, so every warp runs one
//                              attempt_step<S> instantiation on trajectories that behave alike;
//                              between epochs a trajectory is a 568-byte Parked record in HBM
//   dyns_kernel / jac_kernel   stage evaluation only (dual / hyper-dual)
// Compiled with -fmad=false: every fused multiply-add is an explicit fma() (arithmetic contract,
// see radau_core.cuh).
#pragma once
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/janus_b200.h"
#include "radau_core.cuh"
#include "dopri853.cuh"
#include "seulex.cuh"

namespace jb {

int fail(int code, const std::string& msg);   // capi.cu: stores the message for jb200_last_error()
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
#define JB_MINB_FIXED 3   // resident CTAs/SM of the fused fixed-order kernels (168 registers; with H(y0)-ordered rows the
                          // extra warps pay: 55.7 -> 49.5 ms on config 2; the epoch kernels are faster with 2, measured)
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
  double* tout;    // [M, ntspan] (mode 3) / [M, cap] (mode 1, 2) or nullptr
  double* yout;    // [M, ntspan or cap, NY]
  int* nout;       // [M]
  int mode = 3, cap = 0, refine = 1;   // radau_core.cuh: DenseOut
  double* te = nullptr;   // [M, ecap] located events (ev_K > 0)
  double* ye = nullptr;   // [M, ecap, NY]
  int* ie = nullptr;      // [M, ecap]
  int* nev = nullptr;     // [M], zeroed by the host
  int ecap = 0;
  // forward sensitivities through the epochs (epoch_sens_kernel): S [M, NY*NY] is the caller's result matrix, used as the
  // working copy between epochs; sens_cnt [M, 2] = { sensitivity Newton iterations, unconverged (direction, step) pairs }
  double* sens_S = nullptr;
  int* sens_cnt = nullptr;
  double* sens_Sp = nullptr;   // [M, sens_npar, NY] parameter directions (RadauParams::sens_par), or nullptr
};
template <int NY>
__device__ __forceinline__ DenseOut make_dout(const DenseArgs& d, const double* tspan, long long m) {
  const long long rows = d.mode == 3 ? d.ntspan : d.cap;
  DenseOut o{tspan ? tspan + m * d.ntspan : nullptr, d.ntspan, d.tout ? d.tout + m * rows : nullptr,
             d.yout ? d.yout + m * rows * NY : nullptr};
  o.mode = d.mode; o.cap = d.cap; o.refine = d.refine;
  o.te = d.te ? d.te + m * d.ecap : nullptr;
  o.ye = d.ye ? d.ye + m * (long long)d.ecap * NY : nullptr;
  o.ie = d.ie ? d.ie + m * d.ecap : nullptr;
  o.nev = d.nev ? d.nev + m : nullptr;
  o.ecap = d.ecap;
  return o;
}

// CTA shape of the fused kernels.  Fixed 1 / 3 stages: ONE 384-thread CTA per SM (12 warps at 168 registers; measured on config 2,
// fixed s = 3, profiles/r02_fixed_cta_ab.txt: 3 x 128 threads 48.4 ms, 1 x 384 47.3 ms, 2 x 192 48.4, 1 x 256 54.3); the others 2 x 128.
template <int SM, bool VAR>
struct FusedCfg {
#ifdef JB_FUSED_CLASSIC
  static constexpr int threads = JB_THREADS;
  static constexpr int minb = VAR ? JB_MINB_VAR : (SM <= 3 ? JB_MINB_FIXED : 2);
#else
  static constexpr int threads = (!VAR && SM <= 3) ? 384 : JB_THREADS;
  static constexpr int minb = (!VAR && SM <= 3) ? 1 : (VAR ? JB_MINB_VAR : 2);
#endif
};

template <class Ham, int SM, bool VAR, bool DENSE>
__global__ void __launch_bounds__(FusedCfg<SM, VAR>::threads, FusedCfg<SM, VAR>::minb) radau_kernel(const __grid_constant__ RadauParams P,
                                                    const double* __restrict__ y0,
                                                    const double* __restrict__ tspan, long long M,
                                                    double* __restrict__ y_out, double* __restrict__ t_out,
                                                    double* __restrict__ h_out, int* __restrict__ res,
                                                    long long* __restrict__ stats, int* __restrict__ nstg,
                                                    long long* __restrict__ work, const DenseArgs dense,
                                                    unsigned long long* __restrict__ counter,
                                                    const int* __restrict__ order) {
  constexpr int NY = Ham::NV;
  Traj<NY, SM> T;
  long long m = -1;
  bool have = false;
  const int nts = dense.ntspan;
  for (;;) {
    if (!have) {
      m = (long long)atomicAdd(counter, 1ull);
      if (m >= M) break;
      if (order) m = order[m];   // trajectories pre-sorted by H(y0): the lanes of a warp start as phase-space neighbours
      double y[NY];
      load_row<NY>(y0, m, y);
      traj_init<Ham, NY, SM>(T, P, y, __ldg(tspan + m * nts), __ldg(tspan + m * nts + nts - 1));
      have = true;
    }
    const DenseOut dout = DENSE ? make_dout<NY>(dense, tspan, m) : DenseOut{nullptr, 2, nullptr, nullptr};
    const DenseOut* dp = DENSE ? &dout : nullptr;   // compile-time: no cost when no dense output is requested
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
// `fine` (RadauParams::key_fine, epoch_horder 3): 5 more bits = the leading mantissa bits of |h| appended below the
// 10-bit key, so that within a regime the rows are ordered by their CURRENT step size -- an ordering that follows the
// state of the integration instead of H(y0), for batches whose initial states are not a one-parameter family.
constexpr int KEY_FINE_BITS = 5;
__device__ __forceinline__ unsigned short regime_key(int NbrStg, double h, int Newt, int fine = 0) {
  const int k = (NbrStg - 1) >> 1;
  const long long hb = __double_as_longlong(h);
  int e = (int)((hb >> 52) & 0x7ff) - 1023 + 44;   // |h| in [2^-44, 2^20) -> 0..63
  e = e < 0 ? 0 : (e > 63 ? 63 : e);
  int nw = Newt - 1;
  nw = nw < 0 ? 0 : (nw > 3 ? 3 : nw);
  const int coarse = k * 256 + e * 4 + nw;
  if (!fine) return (unsigned short)(1 + coarse);
  const int mant = (int)((hb >> (52 - KEY_FINE_BITS)) & ((1 << KEY_FINE_BITS) - 1));
  return (unsigned short)(1 + ((coarse << KEY_FINE_BITS) | mant));
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
    tag[m] = regime_key(T.NbrStg, T.h, T.Newt, P.key_fine);
  }
}

static __global__ void hist_kernel(const unsigned short* __restrict__ tag, long long M, int* __restrict__ counts, int shift) {
  __shared__ int sh[NKEYS];
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const int t = tag[m];
    if (t) atomicAdd(&sh[(t - 1) >> shift], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x)
    if (sh[i]) atomicAdd(&counts[i], sh[i]);
}

// counting-sort scatter: cursor[key] starts at the exclusive prefix sum of the histogram.
// Warp-aggregated: lanes holding the same key are matched and claim their slots with ONE atomic
// (the live keys are few, so per-lane atomics serialise on a handful of addresses).
static __global__ void scatter_kernel(const unsigned short* __restrict__ tag, long long M, int* __restrict__ cursor,
                               int* __restrict__ list, int shift) {
  const int lane = threadIdx.x & 31;
  for (long long base = blockIdx.x * (long long)blockDim.x; base < M; base += (long long)gridDim.x * blockDim.x) {
    const long long m = base + threadIdx.x;
    const int t = (m < M && tag[m]) ? (((int)tag[m] - 1) >> shift) + 1 : 0;
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

// Epoch kernels: ONE 256-thread CTA per SM (8 warps, the register file's limit at 255 registers per thread).  Measured on
// config 2 (profiles/r02_epoch_cta_ab.txt): 2 x 128 threads 90.4 ms per solve, 1 x 256 for the 5/7-stage kernels 88.9, for all
// epoch kernels 88.3; fewer warps (1 x 128, 2 x 64) 120 ms.
#ifndef JB_THREADS_EPOCH
#define JB_THREADS_EPOCH 256
#endif
#ifndef JB_MINB_EPOCH
#define JB_MINB_EPOCH 1   // resident CTAs/SM for the 1- and 3-stage epoch kernels
#endif
#ifndef JB_MINB_HI
#define JB_MINB_HI 1      // resident CTAs/SM for the 5- and 7-stage epoch kernels (state spills to local memory)
#endif
#ifndef JB_THREADS_HI
#define JB_THREADS_HI 256
#endif
// ---- pre-sort of a batch by the value of the Hamiltonian at the initial state (a scheduling hint only).
// Trajectories that start as neighbours in phase space take nearly the same step sequence, so a warp built from
// 32 neighbours iterates Newton the same number of times, accepts / rejects together and finishes together.
__device__ __forceinline__ unsigned long long ordered_bits(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long u) {
  const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)b);
}
template <class Ham>
__global__ void hkey_kernel(const __grid_constant__ RadauParams P, const double* __restrict__ y0, long long M,
                            unsigned long long* __restrict__ key, int* __restrict__ idx) {
  constexpr int NY = Ham::NV;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    double y[NY];
    load_row<NY>(y0, m, y);
    const double u = Ham::HAS_CONTROL ? Ham::control(y, P.ham_par) : 0.0;
    double v = ham_eval1<Ham>(y, P.ham_par, u, std::make_index_sequence<NY>{}).v;
    if (!(v - v == 0.0)) v = 0.0;
    key[m] = ordered_bits(v);
    idx[m] = (int)m;
  }
}

template <int S>
struct EpochCfg {
  static constexpr int threads = S >= 5 ? JB_THREADS_HI : JB_THREADS_EPOCH;
  static constexpr int minb = S >= 5 ? JB_MINB_HI : JB_MINB_EPOCH;
};

template <class Ham, int S, bool VAR, bool DENSE>
__global__ void __launch_bounds__(EpochCfg<S>::threads, EpochCfg<S>::minb)
    epoch_kernel(const __grid_constant__ RadauParams P, Parked<Ham::NV>* __restrict__ st, const int* __restrict__ list,
                 int count, int max_passes, unsigned long long* __restrict__ counter, unsigned short* __restrict__ tag,
                 double* __restrict__ y_out, double* __restrict__ t_out, double* __restrict__ h_out,
                 int* __restrict__ res, long long* __restrict__ stats, int* __restrict__ nstg,
                 long long* __restrict__ work, const double* __restrict__ tspan, const DenseArgs dense,
                 const int stop_on_accept) {
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
    const DenseOut dout = DENSE ? make_dout<NY>(dense, tspan, m) : DenseOut{nullptr, 2, nullptr, nullptr};
    const DenseOut* dp = DENSE ? &dout : nullptr;
    bool stepped = false;
    const bool fin = run_epoch<Ham, NY, S, VAR>(T, P, head_done, max_passes, dp, stop_on_accept != 0, &stepped);
    park<NY, S>(T, t0, head_done, fin, st[m]);
    tag[m] = (fin || stepped) ? (unsigned short)0 : regime_key(T.NbrStg, T.h, T.Newt, P.key_fine);
    if (fin && y_out) {
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

#ifndef JB_SLIM
// The same epoch with the Ny x Ny sensitivity matrix riding along (radau_core.cuh: sens_step is a hook of attempt_step): S lives
// in the caller's result array between epochs, the real part of the integration is the plain epoch kernel's, bit for bit.
template <class Ham, int S, bool VAR>
__global__ void __launch_bounds__(EpochCfg<S>::threads, EpochCfg<S>::minb)
    epoch_sens_kernel(const __grid_constant__ RadauParams P, Parked<Ham::NV>* __restrict__ st, const int* __restrict__ list,
                      int count, int max_passes, unsigned long long* __restrict__ counter, unsigned short* __restrict__ tag,
                      double* __restrict__ y_out, double* __restrict__ t_out, double* __restrict__ h_out,
                      int* __restrict__ res, long long* __restrict__ stats, int* __restrict__ nstg,
                      long long* __restrict__ work, double* __restrict__ sens_S, int* __restrict__ sens_cnt,
                      double* __restrict__ sens_Sp) {
  constexpr int NY = Ham::NV;
  const int npy = P.sens_npar * NY;
  for (;;) {
    const long long i = (long long)atomicAdd(counter, 1ull);
    if (i >= count) break;
    const long long m = list[i];
    Traj<NY, S> T;
    double t0;
    bool head_done;
    unpark<NY, S>(st[m], T, t0, head_done);
    Sens<NY> sn;
#pragma unroll
    for (int k = 0; k < NY * NY; k++) sn.S[k] = sens_S[m * NY * NY + k];
    for (int k = 0; k < npy; k++) sn.Sp[k] = sens_Sp[m * npy + k];
    sn.iters = sens_cnt[2 * m];
    sn.unconverged = sens_cnt[2 * m + 1];
    const bool fin = run_epoch<Ham, NY, S, VAR>(T, P, head_done, max_passes, nullptr, false, nullptr, &sn);
    park<NY, S>(T, t0, head_done, fin, st[m]);
#pragma unroll
    for (int k = 0; k < NY * NY; k++) sens_S[m * NY * NY + k] = sn.S[k];
    for (int k = 0; k < npy; k++) sens_Sp[m * npy + k] = sn.Sp[k];
    sens_cnt[2 * m] = sn.iters;
    sens_cnt[2 * m + 1] = sn.unconverged;
    tag[m] = fin ? (unsigned short)0 : regime_key(T.NbrStg, T.h, T.Newt, P.key_fine);
    if (fin && y_out) {
      store_row<NY>(y_out, m, T.y);
      t_out[m] = T.t;
      h_out[m] = T.h;
      res[m] = (T.res == 0 && sn.unconverged > 0) ? JB200_RES_SENS_UNCONVERGED : T.res;   // the state is fine, S is not
#pragma unroll
      for (int k = 0; k < 8; k++) stats[m * 8 + k] = (long long)T.st[k];
      if (nstg) nstg[m] = T.NbrStg;
      if (work) work[m] = T.work;
    }
  }
}
#endif

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

// evalDynsDual / evalJacDual: value + dual parts of the dynamics and the Jacobian from one third-order jet of H
struct JacDualArgs {
  long long M;
  int ND;                // dual directions (0: none)
  const double* y;       // [M,NY]
  const double* yd;      // [M,NY,ND] or nullptr
  double *f, *J;         // [M,NY], [M,NY,NY]
  double *T, *fd, *Jd;   // [M,NY,NY,NY], [M,NY,ND], [M,NY,NY,ND]   (each may be nullptr)
};
template <class Ham>
__global__ void __launch_bounds__(128) jac_dual_kernel(const __grid_constant__ RadauParams P, const JacDualArgs a) {
  constexpr int NY = Ham::NV;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < a.M; m += (long long)gridDim.x * blockDim.x) {
    double y[NY], f[NY], J[NY * NY];
    load_row<NY>(a.y, m, y);
    ham_jac_dual<Ham>(y, P.ham_par, P.sign, a.yd ? a.yd + m * NY * a.ND : nullptr, a.ND, f, J,
                      a.T ? a.T + m * NY * NY * NY : nullptr, a.fd ? a.fd + m * NY * a.ND : nullptr,
                      a.Jd ? a.Jd + m * NY * NY * a.ND : nullptr);
    store_row<NY>(a.f, m, f);
#pragma unroll
    for (int r = 0; r < NY; r++) store_row<NY>(a.J, m * NY + r, &J[r * NY]);
  }
}
// ------------------------------------------------------------------ host side
// parked trajectories + the index lists of the epoch scheme; owned by a handle (solve) or by an
// incremental-stepping state (jb200_state_*)
struct EpochWS {
  long long cap = 0, M = 0;
  size_t rec_bytes = 0;
  void* parked = nullptr;
  unsigned short* tag = nullptr;
  int* lists = nullptr;    // [4][cap]: one region per stage count when `order` is in use, else one list [cap]
  int* counts = nullptr;
  int* order = nullptr;    // [cap] rows sorted by H(y0) (phase-space neighbours adjacent)
  bool use_order = false;  // epochs take their lists from `order` (else: regime-key counting sort)
  bool have_order = false;
  int* nsel = nullptr;     // [4] device counts of the per-stage selections
  void* sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int* sens_cnt = nullptr;   // [cap, 2] counters of the sensitivity iteration (epoch_sens_kernel), allocated on first use
  int key_fine = 0;          // this batch's epoch lists are also ordered by the current step size (regime_key `fine`)
};
// the handle's parameters with this workspace's scheduling choice
inline RadauParams params_for(const RadauParams& P, const EpochWS& ws) {
  RadauParams Q = P;
  Q.key_fine = ws.key_fine;
  return Q;
}
inline void ws_free(EpochWS& ws) {
  cudaFree(ws.parked); cudaFree(ws.tag); cudaFree(ws.lists); cudaFree(ws.counts);
  cudaFree(ws.order); cudaFree(ws.nsel); cudaFree(ws.sort_tmp); cudaFree(ws.sens_cnt);
  ws = EpochWS();   // every pointer nulled, cap = 0
}
inline int ws_reserve(EpochWS& ws, long long M, size_t parked_bytes) {
  if (M <= ws.cap && parked_bytes <= ws.rec_bytes) {
    ws.M = M;
    return 0;
  }
  ws_free(ws);      // a failed allocation below leaves nothing dangling
  struct Guard {
    EpochWS& ws; bool ok = false;
    ~Guard() { if (!ok) { ws_free(ws); cudaGetLastError(); } }
  } guard{ws};
  ws.rec_bytes = parked_bytes;
  JB_CUDA(cudaMalloc(&ws.parked, parked_bytes * M));
  JB_CUDA(cudaMalloc(&ws.tag, sizeof(unsigned short) * M));
  JB_CUDA(cudaMalloc(&ws.lists, sizeof(int) * 4 * M));
  JB_CUDA(cudaMalloc(&ws.order, sizeof(int) * M));
  JB_CUDA(cudaMalloc(&ws.nsel, sizeof(int) * 4));
  ws.sort_tmp_bytes = 48 * (size_t)M + (1 << 20);
  JB_CUDA(cudaMalloc(&ws.sort_tmp, ws.sort_tmp_bytes));
  JB_CUDA(cudaMalloc(&ws.counts, sizeof(int) * NKEYS));
  ws.cap = M;
  ws.M = M;
  guard.ok = true;
  return 0;
}

struct Handle {
  int ham_id = 0, ny = 0, device = 0, sm_count = 148;
  RadauParams P;
  int kernel_sm = 3;     // SM template value in use
  bool kernel_var = false;
  unsigned long long* d_counter = nullptr;   // [4] work-fetch counters (one per stage count)
  int launches = 0;
  // variable-order epoch workspace
  EpochWS ws;
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // the per-order epoch kernels of one epoch overlap
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  bool fixed_epochs = false;
  bool presort = true;     // fused fixed-order kernel: batch pre-sorted by H(y0) (JB200_PRESORT=0 disables)
  int epoch_horder = 4;    // epoch lists: 0 regime-key sort, 1 stage count then H(y0) order, 2 regime key then H(y0) order,
                           // 3 regime key then current step size, 4 (default) = 2 or 3 chosen per batch (ws_init: coherence_kernel)
  int epoch_passes = 12;   // attempted steps per trajectory per epoch (flat optimum 8..24 on B200)
  // workspace of jb200_solve_host
  long long cap = 0;
  double *d_y0 = nullptr, *d_ts = nullptr, *d_y = nullptr, *d_t = nullptr, *d_h = nullptr;
  int *d_res = nullptr, *d_nstg = nullptr, *d_stats32 = nullptr;
  long long* d_stats = nullptr;
  long long* d_work = nullptr;
  cudaStream_t own_stream = nullptr;
  void* ext = nullptr;   // per-Hamiltonian context (MLP Hamiltonian: device weights + phase workspace, tu_mlp.cu)
};

// dynamic shared memory of a kernel whose stepper parks complex LU blocks there (lu.cuh: ZLUS)
template <int NY, int S>
inline size_t lc_smem_bytes(int threads) {
  // S = the largest stage count the kernel may run: the unrolled steppers park LcPlace::NS blocks (ZLUS layout), the rolled
  // 5/7-stage steppers all (S-1)/2 blocks in the compact layout (ZLUC); an all-orders kernel needs the larger of the two
  size_t words = (size_t)LcPlace<NY, S>::NS * (size_t)ZLUS<NY>::WORDS;
  if (HiRoll<NY, S>::on) words = std::max(words, (size_t)((S - 1) / 2) * (size_t)ZLUC<NY>::WORDS);
  return sizeof(double) * words * (size_t)threads;
}

// rows sorted by H(y0): radix sort (CUB) of the order-preserving bit pattern of H, payload = row index
template <class Ham>
int build_order(Handle* H, EpochWS& ws, const double* y0, long long M, cudaStream_t st, bool& ok) {
  ok = false;
  const long long gblocks = std::max<long long>(1, std::min<long long>((M + 255) / 256, (long long)H->sm_count * 8));
  unsigned long long* kin = (unsigned long long*)ws.sort_tmp;
  unsigned long long* kout = kin + M;
  int* vin = (int*)(kout + M);
  void* tmp = (void*)(vin + M + (M & 1));
  const size_t tmp_bytes = ws.sort_tmp_bytes - (size_t)((char*)tmp - (char*)ws.sort_tmp);
  size_t need_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need_tmp, kin, kout, vin, ws.order, (int)M, 0, 64, st);
  if (need_tmp > tmp_bytes) return 0;
  hkey_kernel<Ham><<<(unsigned)gblocks, 256, 0, st>>>(H->P, y0, M, kin, vin);
  JB_CUDA(cudaGetLastError());
  JB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, need_tmp, kin, kout, vin, ws.order, (int)M, 0, 64, st));
  ok = true;
  return 0;
}

static __global__ void iota_kernel(int* __restrict__ idx, long long M) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) idx[m] = (int)m;
}

static __global__ void gather_tag_kernel(const unsigned short* __restrict__ tag, const int* __restrict__ order, long long M,
                                         unsigned short* __restrict__ keys) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x)
    keys[m] = tag[order[m]];
}

// live rows of one stage-count bucket, taken in H-order (stable selection)
struct StagePred {
  const unsigned short* tag;
  int k, shift;
  __device__ __forceinline__ bool operator()(const int& idx) const {
    const int t = tag[idx];
    return t != 0 && ((t - 1) >> (8 + shift)) == k;
  }
};

template <class Ham, int SM, bool VAR>
int launch_radau(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                        double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense, cudaStream_t st) {
#ifdef JB_SLIM   // slim plugin builds (large user Hamiltonians): no dense-output kernel variants
  if (dense.nout) return fail(JB200_ENOTSUP, "this plugin was built without dense-output kernels");
  auto kern = radau_kernel<Ham, SM, VAR, false>;
#else
  auto kern = dense.nout ? radau_kernel<Ham, SM, VAR, true> : radau_kernel<Ham, SM, VAR, false>;
#endif
  int bps = 0;
  constexpr int TH = FusedCfg<SM, VAR>::threads;
  const size_t smem = lc_smem_bytes<Ham::NV, SM>(TH);
  if (smem > 48 * 1024) JB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  JB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TH, smem));
  if (bps < 1) bps = 1;
  long long blocks = (long long)H->sm_count * bps;
  const long long need = (M + TH - 1) / TH;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  JB_CUDA(cudaMemsetAsync(H->d_counter, 0, sizeof(unsigned long long), st));
  // pre-sort by H(y0) (large batches)
  const int* order = nullptr;
  if (H->presort && M >= (1 << 16) && M <= 2000000000LL) {
    int rc = ws_reserve(H->ws, M, sizeof(double));
    if (rc) return rc;
    bool ok = false;
    if ((rc = build_order<Ham>(H, H->ws, y0, M, st, ok))) return rc;
    if (ok) order = H->ws.order;
  }
  kern<<<(unsigned)blocks, TH, smem, st>>>(H->P, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense,
                                          H->d_counter, order);
  JB_CUDA(cudaGetLastError());
  H->launches = order ? 2 : 1;   // + the CUB radix-sort passes (library kernels)
  return 0;
}

// outputs of a solve (device arrays; y_out == nullptr: none, incremental stepping)
struct OutArgs {
  double *y_out, *t_out, *h_out;
  int* res;
  long long* stats;
  int* nstg;
  long long* work;
};

template <class Ham, int S, bool VAR>
int launch_epoch(Handle* H, EpochWS& ws, int k, long long base, int count, int max_passes, int stop_on_accept,
                 const OutArgs& o, const double* tspan, const DenseArgs& dense, cudaStream_t main_st, int& used) {
  // first non-empty bucket runs on the caller's stream, the others fork onto auxiliary streams -- unless the list is large
  // enough to fill the GPU by itself (JB200_SERIAL_ROWS, developer A/B): then it queues behind the others on the main stream,
  // so that the kernels of different stage counts (67 / 96 / 134 KB of code) do not share SMs and instruction caches
  static const long long serial_rows = std::getenv("JB200_SERIAL_ROWS") ? atoll(std::getenv("JB200_SERIAL_ROWS")) : 0;
  const bool fork = used > 0 && !(serial_rows > 0 && count >= serial_rows);
  cudaStream_t st = main_st;
  if (fork) {
    st = H->aux[used - 1];
    JB_CUDA(cudaStreamWaitEvent(st, H->ev_fork, 0));
  }
#ifdef JB_SLIM
  if (dense.nout) return fail(JB200_ENOTSUP, "this plugin was built without dense-output kernels");
  auto kern = epoch_kernel<Ham, S, VAR, false>;
#else
  auto kern = dense.nout ? epoch_kernel<Ham, S, VAR, true> : epoch_kernel<Ham, S, VAR, false>;
#endif
  constexpr int TH = EpochCfg<S>::threads;
  int bps = 0;
  const size_t smem = lc_smem_bytes<Ham::NV, S>(TH);
#ifndef JB_SLIM
  if constexpr (VAR) if (dense.sens_S) {
    auto ks = epoch_sens_kernel<Ham, S, VAR>;
    if (smem > 48 * 1024) JB_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ks, TH, smem));
    if (bps < 1) bps = 1;
    long long blocks = (long long)H->sm_count * bps;
    const long long need = ((long long)count + TH - 1) / TH;
    if (blocks > need) blocks = need;
    ks<<<(unsigned)blocks, TH, smem, st>>>(params_for(H->P, ws), (Parked<Ham::NV>*)ws.parked, ws.lists + base, count, max_passes,
                                          H->d_counter + k, ws.tag, o.y_out, o.t_out, o.h_out, o.res, o.stats, o.nstg,
                                          o.work, dense.sens_S, dense.sens_cnt, dense.sens_Sp);
    JB_CUDA(cudaGetLastError());
    H->launches += 1;
    if (fork) {
      JB_CUDA(cudaEventRecord(H->ev_join[used - 1], st));
      JB_CUDA(cudaStreamWaitEvent(main_st, H->ev_join[used - 1], 0));
    }
    if (fork || used == 0) used += 1;
    return 0;
  }
#endif
  if (smem > 48 * 1024) JB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  JB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TH, smem));
  if (bps < 1) bps = 1;
  if (S >= 5 && bps > EpochCfg<S>::minb) bps = EpochCfg<S>::minb;
  long long blocks = (long long)H->sm_count * bps;
  const long long need = ((long long)count + TH - 1) / TH;
  if (blocks > need) blocks = need;
  kern<<<(unsigned)blocks, TH, smem, st>>>(params_for(H->P, ws), (Parked<Ham::NV>*)ws.parked, ws.lists + base, count, max_passes,
                                        H->d_counter + k, ws.tag, o.y_out, o.t_out, o.h_out, o.res, o.stats, o.nstg,
                                        o.work, tspan, dense, stop_on_accept);
  JB_CUDA(cudaGetLastError());
  H->launches += 1;
  if (fork) {
    JB_CUDA(cudaEventRecord(H->ev_join[used - 1], st));
    JB_CUDA(cudaStreamWaitEvent(main_st, H->ev_join[used - 1], 0));
  }
  if (fork || used == 0) used += 1;
  return 0;
}

// { counting sort of the live trajectories by regime key ; one order-homogeneous epoch kernel per
// non-empty stage-count bucket } until no trajectory is tagged live.  Synchronises the stream once
// per epoch (1024 counters).
template <class Ham, bool VAR>
int run_epochs(Handle* H, EpochWS& ws, int max_passes, int stop_on_accept, const OutArgs& o, const double* tspan,
               const DenseArgs& dense, cudaStream_t st) {
  const long long M = ws.M;
  const long long gblocks = std::max<long long>(1, std::min<long long>((M + 255) / 256, (long long)H->sm_count * 8));
  std::vector<int> hist(NKEYS);
  for (long long epoch = 0; ws.use_order && H->epoch_horder == 1; epoch++) {
    // H-ordered epochs: per stage count, the live rows in the order of H(y0) (stable CUB selection from ws.order)
    JB_CUDA(cudaMemsetAsync(H->d_counter, 0, sizeof(unsigned long long) * 4, st));
    for (int k = 0; k < 4; k++) {
      size_t bytes = ws.sort_tmp_bytes;
      JB_CUDA(cub::DeviceSelect::If(ws.sort_tmp, bytes, ws.order, ws.lists + (size_t)k * M, ws.nsel + k, (int)M,
                                    StagePred{ws.tag, k, ws.key_fine ? KEY_FINE_BITS : 0}, st));
    }
    int cnt[4];
    JB_CUDA(cudaMemcpyAsync(cnt, ws.nsel, sizeof cnt, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    H->launches += 1;
    if (cnt[0] + cnt[1] + cnt[2] + cnt[3] == 0) return 0;
    int rc = 0, used = 0;
    JB_CUDA(cudaEventRecord(H->ev_fork, st));
    if (cnt[3]) rc = launch_epoch<Ham, 7, VAR>(H, ws, 3, 3 * M, cnt[3], max_passes, stop_on_accept, o, tspan, dense, st, used);
    if (!rc && cnt[2]) rc = launch_epoch<Ham, 5, VAR>(H, ws, 2, 2 * M, cnt[2], max_passes, stop_on_accept, o, tspan, dense, st, used);
    if (!rc && cnt[1]) rc = launch_epoch<Ham, 3, VAR>(H, ws, 1, 1 * M, cnt[1], max_passes, stop_on_accept, o, tspan, dense, st, used);
    if (!rc && cnt[0]) rc = launch_epoch<Ham, 1, VAR>(H, ws, 0, 0, cnt[0], max_passes, stop_on_accept, o, tspan, dense, st, used);
    if (rc) return rc;
  }
  // regime key major; minor: H(y0) (stable radix sort of the H-ordered rows, horder 2) or the current step size (the fine
  // key bits, rows taken in index order, horder 3)
  const bool key_then_h = (H->epoch_horder >= 2) && ws.have_order;
  static const bool trace = std::getenv("JB200_EPOCH_TRACE") != nullptr;   // developer aid: per-epoch timeline on stderr
  cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
  if (trace) for (auto& e : tev) cudaEventCreate(&e);
  for (long long epoch = 0;; epoch++) {
    if (trace) cudaEventRecord(tev[0], st);
    JB_CUDA(cudaMemsetAsync(ws.counts, 0, sizeof(int) * NKEYS, st));
    JB_CUDA(cudaMemsetAsync(H->d_counter, 0, sizeof(unsigned long long) * 4, st));
    const int shift = ws.key_fine ? KEY_FINE_BITS : 0;
    hist_kernel<<<(unsigned)gblocks, 256, 0, st>>>(ws.tag, M, ws.counts, shift);
    JB_CUDA(cudaGetLastError());
    JB_CUDA(cudaMemcpyAsync(hist.data(), ws.counts, sizeof(int) * NKEYS, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    int cnt[4] = {0, 0, 0, 0}, base[4] = {0, 0, 0, 0}, run = 0;
    for (int i = 0; i < NKEYS; i++) {
      const int c = hist[i];
      hist[i] = run;  // exclusive prefix sum = scatter cursor
      if ((i & 255) == 0) base[i >> 8] = run;
      cnt[i >> 8] += c;
      run += c;
    }
    H->launches += 1;
    if (run == 0) break;
    if (key_then_h) {
      unsigned short* kin = (unsigned short*)ws.sort_tmp;
      unsigned short* kout = kin + M + (M & 1);
      void* tmp = (void*)(kout + M + (M & 1) + 8);
      size_t bytes = ws.sort_tmp_bytes - (size_t)((char*)tmp - (char*)ws.sort_tmp);
      gather_tag_kernel<<<(unsigned)gblocks, 256, 0, st>>>(ws.tag, ws.order, M, kin);
      JB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, ws.order, ws.lists, (int)M, 0, 11 + shift, st));
      const int n0 = (int)(M - run);   // finished rows (key 0) sort to the front
      for (int k = 0; k < 4; k++) base[k] += n0;
    } else {
      JB_CUDA(cudaMemcpyAsync(ws.counts, hist.data(), sizeof(int) * NKEYS, cudaMemcpyHostToDevice, st));
      scatter_kernel<<<(unsigned)gblocks, 256, 0, st>>>(ws.tag, M, ws.counts, ws.lists, shift);
    }
    JB_CUDA(cudaGetLastError());
    H->launches += 1;
    int rc = 0, used = 0;
    if (trace) cudaEventRecord(tev[1], st);
    JB_CUDA(cudaEventRecord(H->ev_fork, st));
    // Stragglers: once the live set no longer fills the GPU (one thread per trajectory), an epoch boundary buys no
    // regrouping, only a sort + a host round trip -- let every trajectory run until it finishes or changes order.
    const int mp = (run < H->sm_count * 2 * JB_THREADS) ? (1 << 30) : max_passes;
    // largest stage count first: its kernel is the longest-running per trajectory
    if (cnt[3]) rc = launch_epoch<Ham, 7, VAR>(H, ws, 3, base[3], cnt[3], mp, stop_on_accept, o, tspan, dense, st, used);
    if (!rc && cnt[2]) rc = launch_epoch<Ham, 5, VAR>(H, ws, 2, base[2], cnt[2], mp, stop_on_accept, o, tspan, dense, st, used);
    if (!rc && cnt[1]) rc = launch_epoch<Ham, 3, VAR>(H, ws, 1, base[1], cnt[1], mp, stop_on_accept, o, tspan, dense, st, used);
    if (!rc && cnt[0]) rc = launch_epoch<Ham, 1, VAR>(H, ws, 0, base[0], cnt[0], mp, stop_on_accept, o, tspan, dense, st, used);
    if (rc) return rc;
    if (trace) {
      cudaEventRecord(tev[2], st);
      cudaEventSynchronize(tev[2]);
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, tev[0], tev[1]);
      cudaEventElapsedTime(&b, tev[1], tev[2]);
      fprintf(stderr, "[epoch %lld] live %d  s1 %d s3 %d s5 %d s7 %d  sort %.3f ms  kernels %.3f ms\n", epoch, run, cnt[0], cnt[1],
              cnt[2], cnt[3], a, b);
    }
  }
  if (trace) for (auto& e : tev) cudaEventDestroy(e);
  return 0;
}

// Is the H(y0) order a phase-space order for THIS batch?  Per state component: sum of squared differences between rows that are
// adjacent in the H order (near) and between rows half a batch apart in it (far).  For a one-parameter family of initial
// states (the reference's examples spread one coordinate: radaute_hamiltonian_example.cpp:108-112) every component is a
// smooth function of H, near / far ~ 1 / M^2; for a cloud with more degrees of freedom rows of equal H are unrelated in the
// other directions and near / far = O(1) for at least one component.
template <int NY>
__global__ void coherence_kernel(const double* __restrict__ y0, const int* __restrict__ order, long long M,
                                 double* __restrict__ acc /* [2][NY] */) {
  double nr[NY], fr[NY];
#pragma unroll
  for (int c = 0; c < NY; c++) nr[c] = fr[c] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i + 1 < M; i += (long long)gridDim.x * blockDim.x) {
    double a[NY], b[NY], f[NY];
    load_row<NY>(y0, order[i], a);
    load_row<NY>(y0, order[i + 1], b);
    load_row<NY>(y0, order[(i + M / 2) % M], f);
#pragma unroll
    for (int c = 0; c < NY; c++) {
      const double dn = a[c] - b[c], df = a[c] - f[c];
      if (dn - dn == 0.0 && df - df == 0.0) {
        nr[c] = fma(dn, dn, nr[c]);
        fr[c] = fma(df, df, fr[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NY; c++) {
    for (int off = 16; off >= 1; off >>= 1) {
      nr[c] += __shfl_xor_sync(0xffffffffu, nr[c], off);
      fr[c] += __shfl_xor_sync(0xffffffffu, fr[c], off);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&acc[c], nr[c]);
      atomicAdd(&acc[NY + c], fr[c]);
    }
  }
}

template <class Ham>
int ws_init(Handle* H, EpochWS& ws, const double* y0, const double* tspan, int ntspan, long long M, cudaStream_t st) {
  constexpr int NY = Ham::NV;
  if (M > 2000000000LL) return fail(JB200_EINVAL, "M too large for the epoch path");
  int rc = ws_reserve(ws, M, sizeof(Parked<NY>));
  if (rc) return rc;
  const long long gblocks = std::max<long long>(1, std::min<long long>((M + 255) / 256, (long long)H->sm_count * 8));
  ws.use_order = ws.have_order = false;
  ws.key_fine = H->epoch_horder == 3;
  bool iota = H->epoch_horder == 3 && M >= (1 << 16);
  if (!iota && H->epoch_horder && M >= (1 << 16)) {
    bool ok = false;
    if ((rc = build_order<Ham>(H, ws, y0, M, st, ok))) return rc;
    ws.use_order = ws.have_order = ok;
    if (ok && H->epoch_horder == 4) {
      // adaptive (default): keep the H(y0) order only if it is a phase-space order for this batch, else order the epoch lists
      // by the current step size -- scheduling only, the results do not depend on it
      double* acc = (double*)ws.sort_tmp;   // free again: build_order's sort has been issued on this stream
      double h_acc[2 * NY];
      JB_CUDA(cudaMemsetAsync(acc, 0, sizeof h_acc, st));
      coherence_kernel<NY><<<(unsigned)gblocks, 256, 0, st>>>(y0, ws.order, M, acc);
      JB_CUDA(cudaGetLastError());
      JB_CUDA(cudaMemcpyAsync(h_acc, acc, sizeof h_acc, cudaMemcpyDeviceToHost, st));
      JB_CUDA(cudaStreamSynchronize(st));
      double worst = 0.0;
      for (int c = 0; c < NY; c++)
        if (h_acc[NY + c] > 0.0) worst = std::max(worst, h_acc[c] / h_acc[NY + c]);
      if (worst > 1e-2) {
        ws.key_fine = 1;
        iota = true;
      }
    }
  }
  if (iota) {
    iota_kernel<<<(unsigned)gblocks, 256, 0, st>>>(ws.order, M);
    JB_CUDA(cudaGetLastError());
    ws.use_order = ws.have_order = true;
  }
  init_kernel<Ham><<<(unsigned)gblocks, 256, 0, st>>>(params_for(H->P, ws), y0, tspan, ntspan, M, (Parked<NY>*)ws.parked, ws.tag);
  JB_CUDA(cudaGetLastError());
  return 0;
}

// variable order (and, optionally, fixed order) through the epochs
template <class Ham, bool VAR = true>
int solve_var(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
              double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense,
              cudaStream_t st) {
  int rc = ws_init<Ham>(H, H->ws, y0, tspan, dense.ntspan, M, st);
  if (rc) return rc;
  H->launches = 1;
  const OutArgs o{y_out, t_out, h_out, res, stats, nstg, work};
  return run_epochs<Ham, VAR>(H, H->ws, H->epoch_passes, 0, o, tspan, dense, st);
}

// ---- incremental stepping (the state / step(htry, ...) protocol of the reference's Python steppers,
//      src/python/seulext.py:8-148,158-581): a state is a set of parked trajectories
template <int NY>
__global__ void rearm_kernel(const __grid_constant__ RadauParams P, Parked<NY>* __restrict__ st,
                             unsigned short* __restrict__ tag, long long M, const double* __restrict__ htry) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    Parked<NY>& p = st[m];
    if (p.flags & PF_FINISHED) {
      tag[m] = 0;
      continue;
    }
    if (htry && htry[m] != p.h) {  // a trial step chosen by the caller: like radau.m:645-655
      const double PosNeg = (p.tfinal - p.t0) >= 0 ? 1.0 : -1.0;
      double h = PosNeg * fmin(fabs(htry[m]), fabs(p.tfinal - p.t0));
      unsigned f = p.flags | PF_NEED_LU;
      if ((p.t + h * 1.0001 - p.tfinal) * PosNeg >= 0) {
        h = p.tfinal - p.t;
        f |= PF_LAST;
      } else {
        f &= ~PF_LAST;
      }
      p.h = h;
      p.flags = f;
    }
    tag[m] = regime_key(p.NbrStg, p.h, p.Newt, P.key_fine);
  }
}

template <int NY>
__global__ void state_read_kernel(const Parked<NY>* __restrict__ st, long long M, double* __restrict__ y,
                                  double* __restrict__ t, double* __restrict__ hnext, double* __restrict__ hdid,
                                  int* __restrict__ res, long long* __restrict__ stats, int* __restrict__ nstg,
                                  int* __restrict__ done) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const Parked<NY>& p = st[m];
#pragma unroll
    for (int i = 0; i < NY; i++) y[m * NY + i] = p.y[i];
    t[m] = p.t;
    hnext[m] = p.h;
    hdid[m] = p.h_old;
    res[m] = p.res;
#pragma unroll
    for (int k = 0; k < 8; k++) stats[m * 8 + k] = (long long)p.st[k];
    if (nstg) nstg[m] = p.NbrStg;
    if (done) done[m] = (p.flags & PF_FINISHED) ? 1 : 0;
  }
}

template <class Ham>
int state_step(Handle* H, EpochWS& ws, const double* htry, cudaStream_t st) {
  constexpr int NY = Ham::NV;
  const long long M = ws.M;
  if (M <= 0) return 0;
  const long long gblocks = std::max<long long>(1, std::min<long long>((M + 255) / 256, (long long)H->sm_count * 8));
  rearm_kernel<NY><<<(unsigned)gblocks, 256, 0, st>>>(params_for(H->P, ws), (Parked<NY>*)ws.parked, ws.tag, M, htry);
  JB_CUDA(cudaGetLastError());
  H->launches = 1;
  const OutArgs none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const DenseArgs nod{2, nullptr, nullptr, nullptr};
  // one ACCEPTED step per live trajectory: no pass budget, stop right after the accept
#ifdef JB_SLIM
  return run_epochs<Ham, true>(H, ws, 1 << 30, 1, none, nullptr, nod, st);
#else
  const bool var = H->P.MinNbrStg != H->P.MaxNbrStg;
  return var ? run_epochs<Ham, true>(H, ws, 1 << 30, 1, none, nullptr, nod, st)
             : run_epochs<Ham, false>(H, ws, 1 << 30, 1, none, nullptr, nod, st);
#endif
}

template <class Ham>
int state_read(Handle* H, EpochWS& ws, double* y, double* t, double* hnext, double* hdid, int* res, long long* stats,
               int* nstg, int* done, cudaStream_t st) {
  constexpr int NY = Ham::NV;
  const long long M = ws.M;
  if (M <= 0) return 0;
  const long long gblocks = std::max<long long>(1, std::min<long long>((M + 255) / 256, (long long)H->sm_count * 8));
  state_read_kernel<NY><<<(unsigned)gblocks, 256, 0, st>>>((const Parked<NY>*)ws.parked, M, y, t, hnext, hdid, res,
                                                          stats, nstg, done);
  JB_CUDA(cudaGetLastError());
  return 0;
}

template <class Ham>
int dispatch_order(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                          double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense, cudaStream_t st) {
  const RadauParams& P = H->P;
#ifdef JB_SLIM   // only the fused 3-stage kernel and the VAR epoch kernels exist (VAR = true also covers Min == Max)
  if (P.MinNbrStg == P.MaxNbrStg && P.MaxNbrStg == 3)
    return launch_radau<Ham, 3, false>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
  return solve_var<Ham, true>(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
#else
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
#endif
}

template <class Ham>
int launch_dyns(Handle* H, const double* y, long long M, double* f, cudaStream_t st) {
  const long long blocks = std::min<long long>((M + 127) / 128, (long long)H->sm_count * 16);
  dyns_kernel<Ham><<<(unsigned)std::max<long long>(blocks, 1), 128, 0, st>>>(H->P, y, M, f);
  JB_CUDA(cudaGetLastError());
  return 0;
}
template <class Ham>
int launch_jac(Handle* H, const double* y, long long M, double* J, cudaStream_t st) {
  const long long blocks = std::min<long long>((M + 127) / 128, (long long)H->sm_count * 16);
  jac_kernel<Ham><<<(unsigned)std::max<long long>(blocks, 1), 128, 0, st>>>(H->P, y, M, J);
  JB_CUDA(cudaGetLastError());
  return 0;
}

// one table of entry points per Hamiltonian: what capi.cu dispatches on
// explicit DOPRI853: one call = one accepted step of every trajectory (dopri853.cuh)
template <class Ham, bool DUAL>   // DUAL: its own kernel, so that the plain stepper keeps its small stack frame
__global__ void __launch_bounds__(128) dopri_kernel(const __grid_constant__ RadauParams P, const dp::StepArgs a) {
  constexpr int NY = Ham::NV;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < a.M; m += (long long)gridDim.x * blockDim.x) {
    double y[NY], d[NY];
    load_row<NY>(a.y, m, y);
    load_row<NY>(a.dydx, m, d);
    double x = a.x[m], errold = a.errold[m], hnext = a.hnext[m], hdid;
    int nrej;
#ifndef JB_SLIM
    if constexpr (DUAL) {   // with the dual part: repeat the accepted attempt stage by stage on [Y, Y'] (dopri853.cuh dopri_dual_one)
      double yb[NY], k1[NY];
#pragma unroll
      for (int i = 0; i < NY; i++) { yb[i] = y[i]; k1[i] = d[i]; }
      const double x0 = x;
      dp::dopri_step_one<Ham>(P.ham_par, P.sign, a.norm_M, a.atol, a.rtol, a.aliased != 0, x, y, d, errold, hnext, a.htry[m], hdid,
                              nrej, yb);
      if (x != x0 && hdid == hdid) {
        double Sd[NY * dp::NDMAX];
        for (int k = 0; k < NY * a.nd; k++) Sd[k] = a.yd[m * NY * a.nd + k];
        dp::dopri_dual_one<Ham>(P.ham_par, P.sign, yb, k1, hdid, Sd, a.nd);
        for (int k = 0; k < NY * a.nd; k++) a.yd[m * NY * a.nd + k] = Sd[k];
      }
    } else
#endif
    dp::dopri_step_one<Ham>(P.ham_par, P.sign, a.norm_M, a.atol, a.rtol, a.aliased != 0, x, y, d, errold, hnext, a.htry[m], hdid, nrej);
    store_row<NY>(a.y, m, y);
    store_row<NY>(a.dydx, m, d);
    a.x[m] = x;
    a.errold[m] = errold;
    a.hnext[m] = hnext;
    a.hdid[m] = hdid;
    if (a.nrej) a.nrej[m] = nrej;
  }
}
template <class Ham>
int launch_dopri(Handle* H, const dp::StepArgs& a, cudaStream_t st) {
  const long long blocks = std::max<long long>(1, std::min<long long>((a.M + 127) / 128, (long long)H->sm_count * 8));
#ifndef JB_SLIM
  if (a.yd) dopri_kernel<Ham, true><<<(unsigned)blocks, 128, 0, st>>>(H->P, a);
  else
#else
  if (a.yd) return fail(JB200_ENOTSUP, "this plugin was built without the dual steppers");
#endif
  dopri_kernel<Ham, false><<<(unsigned)blocks, 128, 0, st>>>(H->P, a);
  JB_CUDA(cudaGetLastError());
  H->launches = 1;
  return 0;
}

// SEULEX: one call = one step() of every trajectory (seulex.cuh)
template <class Ham, bool DUAL>
__global__ void __launch_bounds__(128) seulex_kernel(const __grid_constant__ RadauParams P, const sx::StepArgs a) {
  constexpr int NY = Ham::NV;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < a.M; m += (long long)gridDim.x * blockDim.x) {
    double y[NY], yj[NY];
    load_row<NY>(a.y, m, y);
    load_row<NY>(a.yjac, m, yj);
    double x = a.x[m], theta = a.theta[m], hnext = a.hnext[m], hdid;
    int k_targ = a.k_targ[m], flags = a.flags[m], k, nrej;
#ifndef JB_SLIM
    if constexpr (DUAL) {   // with the dual part: replay the accepted attempt on [y, y'] (seulex.cuh seulex_dual_one)
      double ysav[NY], Sd[NY * sx::NDMAX], Sj[NY * sx::NDMAX], h_taken = 0.0;
      bool fresh = false;
#pragma unroll
      for (int i = 0; i < NY; i++) ysav[i] = y[i];
      sx::seulex_step_one<Ham>(P.ham_par, P.sign, a.atol, a.rtol, a.KMAXX, a.quirks, a.htry[m], y, x, theta, k_targ, hnext, hdid,
                               flags, yj, k, nrej, &fresh, &h_taken);
      const long long o = m * NY * a.nd;
      for (int e = 0; e < NY * a.nd; e++) Sd[e] = a.yd[o + e];
      // a Jacobian evaluated during this call was taken at (a state derived from) the start of the step
      for (int e = 0; e < NY * a.nd; e++) Sj[e] = fresh ? Sd[e] : a.sjac[o + e];
      if (nrej <= 200 && h_taken != 0.0) sx::seulex_dual_one<Ham>(P.ham_par, P.sign, ysav, yj, Sj, h_taken, k, Sd, a.nd);
      for (int e = 0; e < NY * a.nd; e++) { a.yd[o + e] = Sd[e]; a.sjac[o + e] = Sj[e]; }
    } else
#endif
    sx::seulex_step_one<Ham>(P.ham_par, P.sign, a.atol, a.rtol, a.KMAXX, a.quirks, a.htry[m], y, x, theta, k_targ, hnext, hdid,
                             flags, yj, k, nrej);
    store_row<NY>(a.y, m, y);
    store_row<NY>(a.yjac, m, yj);
    a.x[m] = x;
    a.theta[m] = theta;
    a.k_targ[m] = k_targ;
    a.hnext[m] = hnext;
    a.hdid[m] = hdid;
    a.flags[m] = flags;
    if (a.k_out) a.k_out[m] = k;
    if (a.nrej) a.nrej[m] = nrej;
  }
}
template <class Ham>
int launch_seulex(Handle* H, const sx::StepArgs& a, cudaStream_t st) {
#ifdef JB_SLIM
  return fail(JB200_ENOTSUP, "this plugin was built without the SEULEX kernel");
#else
  const long long blocks = std::max<long long>(1, std::min<long long>((a.M + 127) / 128, (long long)H->sm_count * 8));
  if (a.yd) seulex_kernel<Ham, true><<<(unsigned)blocks, 128, 0, st>>>(H->P, a);
  else seulex_kernel<Ham, false><<<(unsigned)blocks, 128, 0, st>>>(H->P, a);
  JB_CUDA(cudaGetLastError());
  H->launches = 1;
  return 0;
#endif
}

template <class Ham>
int launch_jac_dual(Handle* H, const JacDualArgs& a, cudaStream_t st) {
#ifdef JB_SLIM
  return fail(JB200_ENOTSUP, "this plugin was built without the third-order kernel");
#else
  const long long blocks = std::max<long long>(1, std::min<long long>((a.M + 127) / 128, (long long)H->sm_count * 16));
  jac_dual_kernel<Ham><<<(unsigned)blocks, 128, 0, st>>>(H->P, a);
  JB_CUDA(cudaGetLastError());
  return 0;
#endif
}

// forward sensitivities (radau_core.cuh: Sens / sens_step): one thread = one trajectory with its Ny x Ny
// sensitivity matrix, all stage counts inside one kernel (order selection at run time)
#ifndef JB_SENS_THREADS
#define JB_SENS_THREADS 128
#define JB_SENS_MINB 2
#endif
template <class Ham, int SM, bool VAR>
__global__ void __launch_bounds__(JB_SENS_THREADS, JB_SENS_MINB) sens_kernel(const __grid_constant__ RadauParams P, const double* __restrict__ y0,
                                                             const double* __restrict__ tspan, const double* __restrict__ S0,
                                                             long long M, double* __restrict__ y_out, double* __restrict__ S_out,
                                                             double* __restrict__ t_out, double* __restrict__ h_out,
                                                             int* __restrict__ res, long long* __restrict__ stats,
                                                             int* __restrict__ nstg, long long* __restrict__ work,
                                                             int* __restrict__ sens_iters, const int* __restrict__ order,
                                                             double* __restrict__ Sp_out) {
  constexpr int NY = Ham::NV;
  const int npy = P.sens_npar * NY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const long long m = order ? order[i] : i;   // rows in H(y0) order: the lanes of a warp are phase-space neighbours
    Traj<NY, SM> T;
    double y[NY];
    load_row<NY>(y0, m, y);
    traj_init<Ham, NY, SM>(T, P, y, tspan[2 * m], tspan[2 * m + 1]);
    Sens<NY> sn;
    sn.iters = 0;
    sn.unconverged = 0;
#pragma unroll
    for (int k = 0; k < NY * NY; k++) sn.S[k] = S0 ? S0[m * NY * NY + k] : ((k / NY == k % NY) ? 1.0 : 0.0);
    for (int k = 0; k < npy; k++) sn.Sp[k] = 0.0;
    integrate<Ham, NY, SM, VAR>(T, P, nullptr, &sn);
    store_row<NY>(y_out, m, T.y);
#pragma unroll
    for (int k = 0; k < NY * NY; k++) S_out[m * NY * NY + k] = sn.S[k];
    for (int k = 0; k < npy; k++) Sp_out[m * npy + k] = sn.Sp[k];
    t_out[m] = T.t;
    h_out[m] = T.h;
    res[m] = T.res;
#pragma unroll
    for (int k = 0; k < 8; k++) stats[m * 8 + k] = T.st[k];
    if (nstg) nstg[m] = T.NbrStg;
    if (work) work[m] = T.work;
    if (sens_iters) sens_iters[m] = sn.iters;
    if (T.res == 0 && sn.unconverged > 0) res[m] = JB200_RES_SENS_UNCONVERGED;   // the state is fine, S is not
  }
}
template <int NY>
__global__ void sens_seed_kernel(const double* __restrict__ S0, long long M, double* __restrict__ S) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M * NY * NY; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % (NY * NY));
    S[i] = S0 ? S0[i] : ((k / NY == k % NY) ? 1.0 : 0.0);
  }
}
template <class Ham>
int launch_sens(Handle* H, const double* y0, const double* tspan, const double* S0, long long M, double* y_out, double* S_out,
                double* t_out, double* h_out, int* res, long long* stats, int* nstg, long long* work, int* sens_iters,
                double* Sp_out, cudaStream_t st) {
#ifdef JB_SLIM
  return fail(JB200_ENOTSUP, "this plugin was built without the sensitivity kernel");
#else
  // fixed 3 stages (the common fixed-order configuration): one fused kernel.  Everything else -- variable order (the
  // reference default), other fixed stage counts -- goes through the order-homogeneous epochs like the plain solve
  // (epoch_sens_kernel), S riding along in the caller's result matrix.
  constexpr int NY = Ham::NV;
  const bool fixed3 = H->P.MinNbrStg == 3 && H->P.MaxNbrStg == 3;
  if (!fixed3) {
    int rc = ws_init<Ham>(H, H->ws, y0, tspan, 2, M, st);
    if (rc) return rc;
    EpochWS& ws = H->ws;
    if (!ws.sens_cnt) JB_CUDA(cudaMalloc(&ws.sens_cnt, sizeof(int) * 2 * ws.cap));
    JB_CUDA(cudaMemsetAsync(ws.sens_cnt, 0, sizeof(int) * 2 * M, st));
    const long long gblocks = std::max<long long>(1, std::min<long long>((M * NY * NY + 255) / 256, (long long)H->sm_count * 16));
    sens_seed_kernel<NY><<<(unsigned)gblocks, 256, 0, st>>>(S0, M, S_out);
    JB_CUDA(cudaGetLastError());
    H->launches = 2;
    DenseArgs d{2, nullptr, nullptr, nullptr};
    d.sens_S = S_out;
    d.sens_cnt = ws.sens_cnt;
    d.sens_Sp = Sp_out;
    if (H->P.sens_npar > 0) JB_CUDA(cudaMemsetAsync(Sp_out, 0, sizeof(double) * (size_t)M * H->P.sens_npar * NY, st));
    const OutArgs o{y_out, t_out, h_out, res, stats, nstg, work};
    if ((rc = run_epochs<Ham, true>(H, ws, H->epoch_passes, 0, o, tspan, d, st))) return rc;
    if (sens_iters) JB_CUDA(cudaMemcpy2DAsync(sens_iters, sizeof(int), ws.sens_cnt, 2 * sizeof(int), sizeof(int), (size_t)M,
                                              cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  auto kern = sens_kernel<Ham, 3, false>;
  const size_t smem = lc_smem_bytes<Ham::NV, 3>(JB_SENS_THREADS);
  if (smem > 48 * 1024) JB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // many small blocks in flight: trajectories differ in length by up to 5x, the block scheduler balances them
  const long long blocks = std::max<long long>(1, (M + JB_SENS_THREADS - 1) / JB_SENS_THREADS);
  const int* order = nullptr;
  if (H->presort && M >= (1 << 16) && M <= 2000000000LL) {
    int rc = ws_reserve(H->ws, M, sizeof(double));
    if (rc) return rc;
    bool ok = false;
    if ((rc = build_order<Ham>(H, H->ws, y0, M, st, ok))) return rc;
    if (ok) order = H->ws.order;
  }
  kern<<<(unsigned)blocks, JB_SENS_THREADS, smem, st>>>(H->P, y0, tspan, S0, M, y_out, S_out, t_out, h_out, res, stats, nstg, work,
                                                  sens_iters, order, Sp_out);
  JB_CUDA(cudaGetLastError());
  H->launches = 1;
  return 0;
#endif
}

struct HamOps {
  int ny, npar;
  int (*solve)(Handle*, const double*, const double*, long long, double*, double*, double*, int*, long long*, int*,
               long long*, const DenseArgs&, cudaStream_t);
  int (*dyns)(Handle*, const double*, long long, double*, cudaStream_t);
  int (*jac)(Handle*, const double*, long long, double*, cudaStream_t);
  int (*st_init)(Handle*, EpochWS&, const double*, const double*, int, long long, cudaStream_t);
  int (*st_step)(Handle*, EpochWS&, const double*, cudaStream_t);
  int (*st_read)(Handle*, EpochWS&, double*, double*, double*, double*, int*, long long*, int*, int*, cudaStream_t);
  int (*create)(Handle*, const double* par, int npar) = nullptr;   // optional: device-resident parameters
  void (*destroy)(Handle*) = nullptr;
  int (*dopri)(Handle*, const dp::StepArgs&, cudaStream_t) = nullptr;   // explicit DOPRI853 step (thread-resident Hamiltonians)
  int (*sens)(Handle*, const double*, const double*, const double*, long long, double*, double*, double*, double*, int*,
              long long*, int*, long long*, int*, double*, cudaStream_t) = nullptr;   // solve with forward sensitivities
  void (*st_free)(EpochWS&) = nullptr;   // optional: the state of the incremental protocol is not a plain EpochWS
  int (*seulex)(Handle*, const sx::StepArgs&, cudaStream_t) = nullptr;   // SEULEX step (thread-resident Hamiltonians)
  int (*jac_dual)(Handle*, const JacDualArgs&, cudaStream_t) = nullptr;  // evalDynsDual / evalJacDual / third-order ladders
};
template <class Ham>
HamOps make_ops() {
  return HamOps{Ham::NV,           Ham::NPAR,         &dispatch_order<Ham>, &launch_dyns<Ham>,
                &launch_jac<Ham>, &ws_init<Ham>,     &state_step<Ham>,     &state_read<Ham>,
                nullptr,          nullptr,           &launch_dopri<Ham>,   &launch_sens<Ham>,
                nullptr,          &launch_seulex<Ham>, &launch_jac_dual<Ham>};
}

}  // namespace jb
