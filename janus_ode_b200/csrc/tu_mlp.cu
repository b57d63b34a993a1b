// MLP-parameterised Hamiltonian (JB200_HAM_MLP, config 4 of BASELINE.json) behind the same C ABI:
// host orchestration of the phase machine (mlp_phase.cuh) around the tile-GEMM stage evaluations
// (mlp_kernels.cuh).  One pass = one attempted step of every active trajectory:
//
//   head_kernel  ->  [mlp_hess_kernel on the trajectories that need a Jacobian]  ->  prep_kernel
//   repeat { mlp_grad_kernel on the (trajectory, stage) rows still iterating ; newton_pre / solve16 / newton_post }
//   [mlp_grad_kernel at y + err ; estrad_b_kernel]      (first-step retry of the error estimate)
//   mlp_grad_kernel at the accepted points (saves the activations the next Jacobian needs) ; accept_b_kernel
//
// Converged / finished trajectories drop out of the row lists (active-set compaction), so the
// GEMM tiles only ever hold rows that still need work.  The host reads 8 counters once per Newton
// iteration (a few dozen 32-byte copies per solve); all state stays in HBM as structure of arrays.
// parameters: [N = 8, width = 128, W1[128][16], b1[128], W2[128][128], b2, W3[128][128], b3, W4[128], b4]
#include "solver_impl.cuh"
#include "mlp_kernels.cuh"
#include "mlp_phase.cuh"
#include "lu16.cuh"

namespace jb {
namespace mlp {

constexpr int NPAR = 2 + WD * NV + WD + 2 * (WD * WD + WD) + WD + 1;
enum { C_ACTIVE = 0, C_JAC, C_ROWS_A, C_TRAJ_A, C_ROWS_B, C_TRAJ_B, C_ERR, C_F0, C_LU, C_N };

struct Work {              // the phase workspace of one batch (a solve, or a state of the incremental protocol)
  long long cap = 0;       // trajectories it holds
  int cap_SM = 0;
  State S{};
  void* pool = nullptr;
  double* d_act = nullptr;
  int *d_rows[2] = {nullptr, nullptr}, *d_trajs[2] = {nullptr, nullptr};
  int *d_jac = nullptr, *d_err = nullptr, *d_f0 = nullptr, *d_lu = nullptr, *d_cnt = nullptr;
  int* h_cnt = nullptr;    // pinned
};
struct Ctx {
  double* d_w = nullptr;   // all weight matrices, both orientations
  Weights W{};
  Work w;                  // workspace of jb200_solve / jb200_eval_jac
};
struct StepObj {           // jb200_state_*: a batch parked between step() calls
  Work w;
  long long M = 0;
};

__device__ __forceinline__ int append(int* counter, int n) { return atomicAdd(counter, n); }

__global__ void init_kernel(const State S, const __grid_constant__ RadauParams P, const double* __restrict__ y0) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < S.M) ph_init(S, P, i, y0);
}
__global__ void init_b_kernel(const State S) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < S.M) ph_init_b(S, i);
}
__global__ void head_kernel(const State S, const __grid_constant__ RadauParams P, int* cnt, int* jac_list, int stop_on_accept) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= S.M) return;
  ph_head(S, P, i, stop_on_accept != 0);
  const int ph = S.sc[i].phase;
  if (ph == PH_WANT_JAC || ph == PH_PREP) atomicAdd(&cnt[C_ACTIVE], 1);   // takes part in this pass
  if (ph == PH_WANT_JAC) jac_list[append(&cnt[C_JAC], 1)] = (int)i;
}
__global__ void rearm_kernel_mlp(const State S, const double* __restrict__ htry) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < S.M) ph_rearm(S, i, htry);
}
__global__ void state_read_kernel_mlp(const State S, double* __restrict__ y, double* __restrict__ t, double* __restrict__ hnext,
                                      double* __restrict__ hdid, int* __restrict__ res, long long* __restrict__ stats,
                                      int* __restrict__ nstg, int* __restrict__ done) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= S.M) return;
  const TrajS s = S.sc[i];
  for (int e = 0; e < MN; e++) y[i * MN + e] = S.y[(size_t)e * S.M + i];
  t[i] = s.t;
  hnext[i] = s.h;
  hdid[i] = s.h_old;
  res[i] = s.res;
  for (int k = 0; k < 8; k++) stats[i * 8 + k] = s.st[k];
  if (nstg) nstg[i] = s.NbrStg;
  if (done) done[i] = s.phase == PH_FINISHED ? 1 : 0;
}
__device__ __forceinline__ void enqueue(const State& S, long long i, int* cnt, int c_rows, int c_traj, int* rows, int* trajs,
                                        int* err_list, int* f0_list) {
  const TrajS& s = S.sc[i];
  if (s.phase == PH_NEWTON) {
    const int n = s.NbrStg;
    const int at = append(&cnt[c_rows], n);
    for (int q = 0; q < n; q++) rows[at + q] = (int)i * 8 + q;
    trajs[append(&cnt[c_traj], 1)] = (int)i;
  } else if (s.phase == PH_ERR_EVAL) {
    err_list[append(&cnt[C_ERR], 1)] = (int)i * 8;
  } else if (s.phase == PH_F0_EVAL) {
    f0_list[append(&cnt[C_F0], 1)] = (int)i * 8;
  }
}
__global__ void __launch_bounds__(128, 4) prep_kernel(const State S, const __grid_constant__ RadauParams P, int* cnt, int* rows, int* trajs,
                            int* lu_list) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= S.M) return;
  const int ph = S.sc[i].phase;
  if (ph != PH_WANT_JAC && ph != PH_PREP) return;
  ph_prep(S, P, i, true);
  if (S.sc[i].flags & PF_LU_PENDING) lu_list[append(&cnt[C_LU], 1)] = (int)i;
  enqueue(S, i, cnt, C_ROWS_A, C_TRAJ_A, rows, trajs, nullptr, nullptr);
}

// DecomRC (radaute_impl.hpp:1502-1625) for the trajectories ph_prep left pending: 16 lanes per
// trajectory, one matrix row per lane (lu16.cuh); block = 256 threads = 16 trajectories
// x if p else y, as bit logic: written as a ternary on `c == l` inside an unrolled loop over c, the compiler turns the
// initialisation of a row into "negate every entry, then add to entry [l]" -- a run-time-indexed store that drags the whole
// row array into local memory (every later access an LDL / STL)
__device__ __forceinline__ double bit_select(bool p, double x, double y) {
  const long long m = -(long long)p;
  return __longlong_as_double((__double_as_longlong(x) & m) | (__double_as_longlong(y) & ~m));
}

template <int S_>
__device__ __forceinline__ void decomrc16(const State& S, long long i, double h, unsigned mask, int l) {
  using Tab = RadauTab<S_>;
  const double rh = 1.0 / h;
  double J[16];
  {
    const double2* src = reinterpret_cast<const double2*>(S.Jac + (size_t)i * 256 + l * 16);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double2 v = src[j];
      J[2 * j] = v.x;
      J[2 * j + 1] = v.y;
    }
  }
  {
    const double valp0 = Tab::ValP(0) * rh;
    double a[16], rd = 0.0;
    int piv = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) a[c] = bit_select(c == l, valp0 - J[c], -J[c]);
    rlu16_factor(mask, l, a, rd, piv);
    double2* dst = reinterpret_cast<double2*>(S.L0a(i) + l * 16);
#pragma unroll
    for (int j = 0; j < 8; j++) dst[j] = make_double2(a[2 * j], a[2 * j + 1]);
    S.L0rd(i)[l] = rd;
    S.L0piv(i)[l] = piv;
  }
#pragma unroll
  for (int q = 1; q <= (S_ - 1) / 2; q++) {
    const double al = Tab::ValP(2 * q - 1) * rh, be = Tab::ValP(2 * q) * rh;
    double re[16], im[16], rdre = 0.0, rdim = 0.0;
    int piv = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) {
      re[c] = bit_select(c == l, al - J[c], -J[c]);
      im[c] = bit_select(c == l, be, 0.0);
    }
    zlu16_factor(mask, l, re, im, rdre, rdim, piv);
    double2* dr = reinterpret_cast<double2*>(S.Lre(i, q - 1) + l * 16);
    double2* di = reinterpret_cast<double2*>(S.Lim(i, q - 1) + l * 16);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      dr[j] = make_double2(re[2 * j], re[2 * j + 1]);
      di[j] = make_double2(im[2 * j], im[2 * j + 1]);
    }
    S.Lrdre(i, q - 1)[l] = rdre;
    S.Lrdim(i, q - 1)[l] = rdim;
    S.Lpiv(i, q - 1)[l] = piv;
  }
}

// One instantiation per stage count (a half-warp whose trajectory has another stage count leaves at once): with the four
// decomrc16<S> bodies behind a switch in ONE kernel the compiler kept the row arrays in local memory (286 STL + 202 LDL in the
// SASS, ~50 kB of local traffic per trajectory); alone, each body keeps its rows in registers.
template <int S_>
__global__ void __launch_bounds__(256) lu16_kernel(const State S, const int* __restrict__ lu_list, int n) {
  const int slot = blockIdx.x * 16 + (threadIdx.x >> 4);
  if (slot >= n) return;   // whole half-warp leaves together
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  const int l = threadIdx.x & 15;
  const long long i = lu_list[slot];
  if (S.sc[i].NbrStg != S_) return;   // uniform over the 16 lanes
  const double h = S.sc[i].h;
  __syncwarp(mask);
  if (l == 0) S.sc[i].flags &= ~PF_LU_PENDING;
  decomrc16<S_>(S, i, h, mask, l);
}
static void launch_lu16(const State& S, const int* lu_list, int n, int min_stg, int max_stg, cudaStream_t st) {
  const unsigned grid = (unsigned)((n + 15) / 16);
  if (min_stg <= 1 && max_stg >= 1) lu16_kernel<1><<<grid, 256, 0, st>>>(S, lu_list, n);
  if (min_stg <= 3 && max_stg >= 3) lu16_kernel<3><<<grid, 256, 0, st>>>(S, lu_list, n);
  if (min_stg <= 5 && max_stg >= 5) lu16_kernel<5><<<grid, 256, 0, st>>>(S, lu_list, n);
  if (min_stg <= 7 && max_stg >= 7) lu16_kernel<7><<<grid, 256, 0, st>>>(S, lu_list, n);
}
// one Newton iteration = newton_pre_kernel (z = TI f, right-hand sides; one thread per trajectory)
//                      -> solve16_kernel (Solvrad's triangular solves; 16 lanes per trajectory, lu16.cuh)
//                      -> newton_post_kernel (norms, convergence logic, Estrad / controller; one thread)
__global__ void __launch_bounds__(128, 4) newton_pre_kernel(const State S, const __grid_constant__ RadauParams P, const int* __restrict__ trajs_in,
                                  int n_in) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_in) return;
  ph_newton_pre(S, P, trajs_in[k]);
}
__global__ void __launch_bounds__(256) solve16_kernel(const State S, const int* __restrict__ trajs_in, int n_in) {
  const int slot = blockIdx.x * 16 + (threadIdx.x >> 4);
  if (slot >= n_in) return;
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  const int l = threadIdx.x & 15;
  const long long i = trajs_in[slot];
  if (S.sc[i].phase != PH_NEWTON) return;   // finished by newton_pre (uniform over the 16 lanes)
  const int nq = (S.sc[i].NbrStg - 1) / 2;
  const long long M = S.M;
  {
    double a[16];
    const double2* src = reinterpret_cast<const double2*>(S.L0a(i) + l * 16);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double2 v = src[j];
      a[2 * j] = v.x;
      a[2 * j + 1] = v.y;
    }
    double b = S.z[(size_t)l * M + i];
    rlu16_solve(mask, l, a, S.L0rd(i)[l], S.L0piv(i)[l], b);
    S.z[(size_t)l * M + i] = b;
  }
  for (int q = 0; q < nq; q++) {
    double re[16], im[16];
    const double2* sr = reinterpret_cast<const double2*>(S.Lre(i, q) + l * 16);
    const double2* si = reinterpret_cast<const double2*>(S.Lim(i, q) + l * 16);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double2 v = sr[j], u = si[j];
      re[2 * j] = v.x;
      re[2 * j + 1] = v.y;
      im[2 * j] = u.x;
      im[2 * j + 1] = u.y;
    }
    double* zr = S.z + (size_t)((2 * q + 1) * MN + l) * M + i;
    double* zi = S.z + (size_t)((2 * q + 2) * MN + l) * M + i;
    double br = *zr, bi = *zi;
    zlu16_solve(mask, l, re, im, S.Lrdre(i, q)[l], S.Lrdim(i, q)[l], S.Lpiv(i, q)[l], br, bi);
    *zr = br;
    *zi = bi;
  }
}
__global__ void __launch_bounds__(128, 4) newton_post_kernel(const State S, const __grid_constant__ RadauParams P, const int* __restrict__ trajs_in,
                                   int n_in, int* cnt, int c_rows, int c_traj, int* rows_out, int* trajs_out, int* err_list,
                                   int* f0_list) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_in) return;
  const long long i = trajs_in[k];
  ph_newton_post(S, P, i);
  enqueue(S, i, cnt, c_rows, c_traj, rows_out, trajs_out, err_list, f0_list);
}
__global__ void estrad_b_kernel(const State S, const __grid_constant__ RadauParams P, const int* __restrict__ err_list, int n,
                                int* cnt, int* f0_list) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long long i = err_list[k] >> 3;
  ph_estrad_b(S, P, i);
  if (S.sc[i].phase == PH_F0_EVAL) f0_list[append(&cnt[C_F0], 1)] = (int)i * 8;
}
__global__ void accept_b_kernel(const State S, const __grid_constant__ RadauParams P, const int* __restrict__ f0_list, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  ph_accept_b(S, P, f0_list[k] >> 3);
}
__global__ void output_kernel(const State S, double* __restrict__ y_out, double* __restrict__ t_out, double* __restrict__ h_out,
                              int* __restrict__ res, long long* __restrict__ stats, int* __restrict__ nstg,
                              long long* __restrict__ work, int* __restrict__ nout) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= S.M) return;
  const TrajS s = S.sc[i];
  if (S.tout) {
    dense_finish_g(S, s, i);
    nout[i] = s.nout;
  }
  for (int e = 0; e < MN; e++) y_out[i * MN + e] = S.y[(size_t)e * S.M + i];
  t_out[i] = s.t;
  h_out[i] = s.h;
  res[i] = s.res;
  for (int k = 0; k < 8; k++) stats[i * 8 + k] = s.st[k];
  if (nstg) nstg[i] = s.NbrStg;
  if (work) work[i] = s.work;
}

// jb200_lu_solve for n = 16: factor + solve of M independent systems, 16 lanes per system
template <bool CPLX>
__global__ void __launch_bounds__(256) lu16_solve_kernel(long long M, const double* __restrict__ Are,
                                                         const double* __restrict__ Aim, const double* __restrict__ bre,
                                                         const double* __restrict__ bim, double* __restrict__ xre,
                                                         double* __restrict__ xim) {
  const long long m = blockIdx.x * 16LL + (threadIdx.x >> 4);
  if (m >= M) return;
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  const int l = threadIdx.x & 15;
  double re[16], im[16];
#pragma unroll
  for (int c = 0; c < 16; c++) {
    re[c] = Are[m * 256 + l * 16 + c];
    im[c] = CPLX ? Aim[m * 256 + l * 16 + c] : 0.0;
  }
  int piv = 0;
  if (CPLX) {
    double rdre = 0.0, rdim = 0.0;
    zlu16_factor(mask, l, re, im, rdre, rdim, piv);
    double br = bre[m * 16 + l], bi = bim[m * 16 + l];
    zlu16_solve(mask, l, re, im, rdre, rdim, piv, br, bi);
    xre[m * 16 + l] = br;
    xim[m * 16 + l] = bi;
  } else {
    double rd = 0.0;
    rlu16_factor(mask, l, re, rd, piv);
    double b = bre[m * 16 + l];
    rlu16_solve(mask, l, re, rd, piv, b);
    xre[m * 16 + l] = b;
  }
}

static Ctx* ctx_of(Handle* H) { return (Ctx*)H->ext; }

static void free_workspace(Work& C) {
  cudaFree(C.pool);
  cudaFree(C.d_act);
  for (int k = 0; k < 2; k++) { cudaFree(C.d_rows[k]); cudaFree(C.d_trajs[k]); }
  cudaFree(C.d_jac); cudaFree(C.d_err); cudaFree(C.d_f0); cudaFree(C.d_lu); cudaFree(C.d_cnt);
  if (C.h_cnt) cudaFreeHost(C.h_cnt);
  C = Work();
}

static int ensure(Handle* H, Work& C, long long M, bool need_state) {
  const int SM = H->P.MaxNbrStg;
  if (M <= C.cap && SM <= C.cap_SM && (C.pool || !need_state)) return 0;
  if (M > (1LL << 27)) return fail(JB200_EINVAL, "MLP Hamiltonian path: M must be <= 2^27 per GPU");
  free_workspace(C);
  JB_CUDA(cudaMalloc(&C.d_act, sizeof(double) * 5 * WD * (size_t)M));
  JB_CUDA(cudaMalloc(&C.d_cnt, sizeof(int) * C_N));
  JB_CUDA(cudaMallocHost(&C.h_cnt, sizeof(int) * C_N));
  if (need_state) {
    const size_t per_d = MN * 5 + MN * MN + (size_t)SM * MN * 4 + 1 + (size_t)lu_doubles(SM);
    const size_t per_i = (size_t)lu_ints(SM);
    const size_t bytes = sizeof(TrajS) * (size_t)M + 1024 + sizeof(double) * per_d * (size_t)M + sizeof(int) * per_i * (size_t)M;
    JB_CUDA(cudaMalloc(&C.pool, bytes));
    char* p = (char*)C.pool;
    State& S = C.S;
    S.M = M;
    S.SM = SM;
    S.sc = (TrajS*)p;
    p += ((sizeof(TrajS) * (size_t)M + 255) / 256) * 256;
    double* d = (double*)p;
    auto take = [&](size_t per) { double* r = d; d += ((per * (size_t)M + 1) / 2) * 2; return r; };   // 16-byte aligned
    S.y = take(MN); S.f0 = take(MN); S.cont = take(SM * MN); S.z = take(SM * MN);
    S.w = take(SM * MN); S.F = take(SM * MN); S.rscal = take(MN); S.temp = take(MN); S.ev = take(MN); S.errs = take(1);
    S.Jac = take(MN * MN);
    S.lud = lu_doubles(SM);
    S.lui = lu_ints(SM);
    S.LU = take(S.lud);
    S.LUpiv = (int*)d;
    for (int k = 0; k < 2; k++) {
      JB_CUDA(cudaMalloc(&C.d_rows[k], sizeof(int) * (size_t)SM * M));
      JB_CUDA(cudaMalloc(&C.d_trajs[k], sizeof(int) * (size_t)M));
    }
    JB_CUDA(cudaMalloc(&C.d_jac, sizeof(int) * (size_t)M));
    JB_CUDA(cudaMalloc(&C.d_err, sizeof(int) * (size_t)M));
    JB_CUDA(cudaMalloc(&C.d_f0, sizeof(int) * (size_t)M));
    JB_CUDA(cudaMalloc(&C.d_lu, sizeof(int) * (size_t)M));
  }
  C.cap = M;
  C.cap_SM = SM;
  return 0;
}

static int launch_grad(Handle* H, const GradIO& io, int n_rows, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  const int tiles = (n_rows + TR - 1) / TR;
  const int grid = tiles < H->sm_count * CTAS_PER_SM ? tiles : H->sm_count * CTAS_PER_SM;
  mlp_grad_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ctx_of(H)->W, io);
  JB_CUDA(cudaGetLastError());
  H->launches += 1;
  return 0;
}
static int launch_hess(Handle* H, const HessIO& io, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  const int tiles = (n + TR / 16 - 1) / (TR / 16);
  const int grid = tiles < H->sm_count * CTAS_PER_SM ? tiles : H->sm_count * CTAS_PER_SM;
  mlp_hess_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ctx_of(H)->W, io);
  JB_CUDA(cudaGetLastError());
  H->launches += 1;
  return 0;
}

static int read_counters(Work* C, cudaStream_t st) {
  JB_CUDA(cudaMemcpyAsync(C->h_cnt, C->d_cnt, sizeof(int) * C_N, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// initial state of every trajectory: ph_init, f(t0, y0) (activations saved for the first Jacobian), NaN check
static int begin(Handle* H, Work* C, State& S, const double* y0, long long M, cudaStream_t st) {
  const unsigned gb = (unsigned)((M + 127) / 128);
  init_kernel<<<gb, 128, 0, st>>>(S, H->P, y0);
  GradIO g0{nullptr, nullptr, (int)M, S.y, 1, M, nullptr, 0, 0, nullptr, S.f0, 1, M, C->d_act, H->P.sign};
  const int rc = launch_grad(H, g0, (int)M, st);
  if (rc) return rc;
  init_b_kernel<<<gb, 128, 0, st>>>(S);
  H->launches += 2;
  JB_CUDA(cudaGetLastError());
  return 0;
}

// passes (one attempted step of every participating trajectory each) until nobody is left; stop_on_accept:
// a trajectory stops participating once it has accepted a step in this call (incremental protocol)
static int run_passes(Handle* H, Work* C, State& S, bool stop_on_accept, cudaStream_t st) {
  const long long M = S.M;
  const RadauParams& P = H->P;
  const unsigned gb = (unsigned)((M + 127) / 128);
  auto blocks = [](int n) { return (unsigned)((n + 127) / 128); };
  int rc = 0;
  for (;;) {
    JB_CUDA(cudaMemsetAsync(C->d_cnt, 0, sizeof(int) * C_N, st));
    head_kernel<<<gb, 128, 0, st>>>(S, P, C->d_cnt, C->d_jac, stop_on_accept ? 1 : 0);
    H->launches += 1;
    if ((rc = read_counters(C, st))) return rc;
    if (C->h_cnt[C_ACTIVE] == 0) break;
    if (C->h_cnt[C_JAC] > 0) {
      HessIO hio{C->d_jac, nullptr, C->h_cnt[C_JAC], C->d_act, S.Jac, MN * MN, 1, P.sign};
      if ((rc = launch_hess(H, hio, C->h_cnt[C_JAC], st))) return rc;
    }
    prep_kernel<<<gb, 128, 0, st>>>(S, P, C->d_cnt, C->d_rows[0], C->d_trajs[0], C->d_lu);
    H->launches += 1;
    int cur = 0;
    for (int it = 0;; it++) {
      if ((rc = read_counters(C, st))) return rc;
      if (it == 0 && C->h_cnt[C_LU] > 0) {
        launch_lu16(S, C->d_lu, C->h_cnt[C_LU], H->P.MinNbrStg, H->P.MaxNbrStg, st);
        H->launches += (H->P.MaxNbrStg - H->P.MinNbrStg) / 2 + 1;
      }
      const int n_rows = C->h_cnt[cur ? C_ROWS_B : C_ROWS_A], n_traj = C->h_cnt[cur ? C_TRAJ_B : C_TRAJ_A];
      if (n_traj == 0) break;
      GradIO gi{C->d_rows[cur], nullptr, n_rows, S.y, 1, M, S.z, 1, M, nullptr, S.F, 1, M, nullptr, P.sign};
      if ((rc = launch_grad(H, gi, n_rows, st))) return rc;
      const int nxt = cur ^ 1;
      JB_CUDA(cudaMemsetAsync(C->d_cnt + (nxt ? C_ROWS_B : C_ROWS_A), 0, 2 * sizeof(int), st));
      newton_pre_kernel<<<blocks(n_traj), 128, 0, st>>>(S, P, C->d_trajs[cur], n_traj);
      solve16_kernel<<<(unsigned)((n_traj + 15) / 16), 256, 0, st>>>(S, C->d_trajs[cur], n_traj);
      newton_post_kernel<<<blocks(n_traj), 128, 0, st>>>(S, P, C->d_trajs[cur], n_traj, C->d_cnt, nxt ? C_ROWS_B : C_ROWS_A,
                                                         nxt ? C_TRAJ_B : C_TRAJ_A, C->d_rows[nxt], C->d_trajs[nxt],
                                                         C->d_err, C->d_f0);
      H->launches += 3;
      cur = nxt;
    }
    if (C->h_cnt[C_ERR] > 0) {
      const int n = C->h_cnt[C_ERR];
      GradIO ge{C->d_err, nullptr, n, S.y, 1, M, nullptr, 0, 0, S.errs, S.ev, 1, M, nullptr, P.sign};
      if ((rc = launch_grad(H, ge, n, st))) return rc;
      estrad_b_kernel<<<blocks(n), 128, 0, st>>>(S, P, C->d_err, n, C->d_cnt, C->d_f0);
      H->launches += 1;
      if ((rc = read_counters(C, st))) return rc;
    }
    if (C->h_cnt[C_F0] > 0) {
      const int n = C->h_cnt[C_F0];
      GradIO gf{C->d_f0, nullptr, n, S.y, 1, M, nullptr, 0, 0, nullptr, S.f0, 1, M, C->d_act, P.sign};
      if ((rc = launch_grad(H, gf, n, st))) return rc;
      accept_b_kernel<<<blocks(n), 128, 0, st>>>(S, P, C->d_f0, n);
      H->launches += 1;
    }
  }
  JB_CUDA(cudaGetLastError());
  return 0;
}

static State bind(Work* C, long long M, const double* tspan, const DenseArgs& dense) {
  State S = C->S;
  S.M = M;   // element stride = this batch (the pool was sized for cap >= M)
  S.ntspan = dense.ntspan;
  S.tspan = tspan;
  S.tout = dense.tout;
  S.yout = dense.yout;
  return S;
}

static int solve(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out, double* h_out,
                 int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense, cudaStream_t st) {
  if (dense.nout && (dense.mode != 3 || dense.nev))
    return fail(JB200_ENOTSUP, "the MLP-Hamiltonian path records output at tspan points only (no per-step output, no events)");
  Work* C = &ctx_of(H)->w;
  int rc = ensure(H, *C, M, true);
  if (rc) return rc;
  State S = bind(C, M, tspan, dense);
  H->launches = 0;
  if ((rc = begin(H, C, S, y0, M, st))) return rc;
  if ((rc = run_passes(H, C, S, false, st))) return rc;
  output_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(S, y_out, t_out, h_out, res, stats, nstg, work, dense.nout);
  H->launches += 1;
  JB_CUDA(cudaGetLastError());
  return 0;
}

static int dyns(Handle* H, const double* y, long long M, double* ydot, cudaStream_t st) {
  const GradIO io{nullptr, nullptr, (int)M, y, MN, 1, nullptr, 0, 0, nullptr, ydot, MN, 1, nullptr, H->P.sign};
  return launch_grad(H, io, (int)M, st);
}

static int jac(Handle* H, const double* y, long long M, double* J, cudaStream_t st) {
  Work* C = &ctx_of(H)->w;
  int rc = ensure(H, *C, M, false);
  if (rc) return rc;
  double* tmp = nullptr;
  JB_CUDA(cudaMalloc(&tmp, sizeof(double) * MN * (size_t)M));
  const GradIO io{nullptr, nullptr, (int)M, y, MN, 1, nullptr, 0, 0, nullptr, tmp, MN, 1, C->d_act, H->P.sign};
  rc = launch_grad(H, io, (int)M, st);
  if (!rc) {
    const HessIO hio{nullptr, nullptr, (int)M, C->d_act, J, MN * MN, 1, H->P.sign};
    rc = launch_hess(H, hio, (int)M, st);
  }
  cudaStreamSynchronize(st);
  cudaFree(tmp);
  return rc;
}

static int create(Handle* H, const double* par, int npar) {
  if (npar != NPAR) return fail(JB200_EINVAL, "MLP Hamiltonian needs " + std::to_string(NPAR) + " parameters");
  if ((int)par[0] != NH || (int)par[1] != WD)
    return fail(JB200_ENOTSUP, "MLP Hamiltonian kernels are built for N = 8 (16 inputs), width 128");
  Ctx* C = new Ctx();
  H->ext = C;
  const double* W1 = par + 2;
  const double* b1 = W1 + WD * NV;
  const double* W2 = b1 + WD;
  const double* b2 = W2 + WD * WD;
  const double* W3 = b2 + WD;
  const double* b3 = W3 + WD * WD;
  const double* W4 = b3 + WD;
  std::vector<double> buf;
  auto push = [&](const double* src, int rows, int cols, bool transpose) {
    const size_t at = buf.size();
    buf.resize(at + (size_t)rows * cols);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        if (transpose) buf[at + (size_t)c * rows + r] = src[(size_t)r * cols + c];
        else buf[at + (size_t)r * cols + c] = src[(size_t)r * cols + c];
      }
    return at;
  };
  const size_t oW1t = push(W1, WD, NV, true), oW2t = push(W2, WD, WD, true), oW3t = push(W3, WD, WD, true);
  const size_t oW3 = push(W3, WD, WD, false), oW2 = push(W2, WD, WD, false), oW1 = push(W1, WD, NV, false);
  const size_t ob1 = push(b1, 1, WD, false), ob2 = push(b2, 1, WD, false), ob3 = push(b3, 1, WD, false);
  const size_t oW4 = push(W4, 1, WD, false);
  JB_CUDA(cudaMalloc(&C->d_w, sizeof(double) * buf.size()));
  JB_CUDA(cudaMemcpy(C->d_w, buf.data(), sizeof(double) * buf.size(), cudaMemcpyHostToDevice));
  C->W = Weights{C->d_w + oW1t, C->d_w + oW2t, C->d_w + oW3t, C->d_w + oW3, C->d_w + oW2, C->d_w + oW1,
                 C->d_w + ob1,  C->d_w + ob2,  C->d_w + ob3,  C->d_w + oW4};
  JB_CUDA(cudaFuncSetAttribute(mlp_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  JB_CUDA(cudaFuncSetAttribute(mlp_hess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  return 0;
}

static void destroy(Handle* H) {
  Ctx* C = ctx_of(H);
  if (!C) return;
  free_workspace(C->w);
  cudaFree(C->d_w);
  delete C;
  H->ext = nullptr;
}

// ---- incremental protocol (jb200_state_*): the batch lives in its own workspace between step() calls; the
// EpochWS of the generic State only carries the pointer (parked) and the batch size
static int st_init(Handle* H, EpochWS& ws, const double* y0, const double* tspan, int ntspan, long long M, cudaStream_t st) {
  if (ntspan != 2) return fail(JB200_EINVAL, "state protocol: tspan must be [M,2]");
  StepObj* o = new StepObj();
  o->M = M;
  ws.parked = o;
  ws.M = M;
  int rc = ensure(H, o->w, M, true);
  if (rc) return rc;
  const DenseArgs none{2, nullptr, nullptr, nullptr};
  // tspan is only read here (t0, tfinal go into the per-trajectory scalars): the caller may free it afterwards
  State S = bind(&o->w, M, tspan, none);
  H->launches = 0;
  return begin(H, &o->w, S, y0, M, st);
}
static int st_step(Handle* H, EpochWS& ws, const double* htry, cudaStream_t st) {
  StepObj* o = (StepObj*)ws.parked;
  if (!o || o->M <= 0) return 0;
  const DenseArgs none{2, nullptr, nullptr, nullptr};
  State S = bind(&o->w, o->M, nullptr, none);
  H->launches = 1;
  rearm_kernel_mlp<<<(unsigned)((o->M + 127) / 128), 128, 0, st>>>(S, htry);
  return run_passes(H, &o->w, S, true, st);
}
static int st_read(Handle*, EpochWS& ws, double* y, double* t, double* hnext, double* hdid, int* res, long long* stats, int* nstg,
                   int* done, cudaStream_t st) {
  StepObj* o = (StepObj*)ws.parked;
  if (!o || o->M <= 0) return 0;
  const DenseArgs none{2, nullptr, nullptr, nullptr};
  State S = bind(&o->w, o->M, nullptr, none);
  state_read_kernel_mlp<<<(unsigned)((o->M + 127) / 128), 128, 0, st>>>(S, y, t, hnext, hdid, res, stats, nstg, done);
  JB_CUDA(cudaGetLastError());
  return 0;
}
static void st_free(EpochWS& ws) {
  StepObj* o = (StepObj*)ws.parked;
  if (o) {
    free_workspace(o->w);
    delete o;
  }
  ws = EpochWS();
}

}  // namespace mlp

int lu16_solve_launch(long long M, const double* Are, const double* Aim, const double* bre, const double* bim, double* xre,
                      double* xim, cudaStream_t st) {
  const unsigned blocks = (unsigned)((M + 15) / 16);
  if (Aim) mlp::lu16_solve_kernel<true><<<blocks, 256, 0, st>>>(M, Are, Aim, bre, bim, xre, xim);
  else mlp::lu16_solve_kernel<false><<<blocks, 256, 0, st>>>(M, Are, nullptr, bre, nullptr, xre, nullptr);
  JB_CUDA(cudaGetLastError());
  return 0;
}

extern const HamOps jb_ops_mlp = HamOps{mlp::MN,       mlp::NPAR,     &mlp::solve,   &mlp::dyns,   &mlp::jac,
                                        &mlp::st_init, &mlp::st_step, &mlp::st_read, &mlp::create, &mlp::destroy,
                                        nullptr,       nullptr,       &mlp::st_free};

}  // namespace jb
