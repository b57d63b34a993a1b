// C ABI (include/janus_b200.h) of the Hamiltonian Radau path, sm_100a: option validation, per-solve
// constants, handle / workspace management and dispatch to the per-Hamiltonian entry points
// (solver_impl.cuh, one translation unit per Hamiltonian: tu_ham.cu), plus the Hamiltonian-independent
// kernels: lu_kernel (batched small dense LU factor+solve, real / complex), detpow_kernel (test hook),
// dfma_kernel (FP64 roofline probe).
//
// Build modes:  default      -- capi.cu + 5 x tu_ham.cu, linked (see __graft_entry__.build())
//               -DJB_USER_HAM_HEADER="file"  monolithic plugin around ONE user Hamiltonian
//                                            `jb::UserHam` (janus_ode_b200/user.py) = JB200_HAM_USER
//               -DJB_FAST_BUILD               monolithic, van der Pol only (developer A/B builds)
#include <future>

#include "solver_impl.cuh"
#ifdef JB_USER_HAM_HEADER
#include JB_USER_HAM_HEADER
#endif

namespace jb {

static thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// ------------------------------------------------------------------ Hamiltonian-independent kernels
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

// FP64 tensor-core probe: independent chains of mma.sync.aligned.m8n8k4.f64 (DMMA) per warp, 512 FLOP each.
// tcgen05 has no f64 kind, so DMMA is the only tensor-core route for an fp64 GEMM on sm_100a; this measures whether it
// beats the DFMA pipe (it would decide the inner loops of the MLP-Hamiltonian GEMMs, csrc/mlp_kernels.cuh).
__global__ void dmma_kernel(int iters, double seed, double* sink) {
  constexpr int NCH = 8;
  double c0[NCH], c1[NCH];
#pragma unroll
  for (int k = 0; k < NCH; k++) {
    c0[k] = seed + k * 1e-3;
    c1[k] = seed - k * 1e-3;
  }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s += c0[k] + c1[k];
  if (s == 123.456) sink[0] = s;
}

// ------------------------------------------------------------------ Hamiltonian table
#if defined(JB_USER_HAM_HEADER)
static const HamOps jb_ops_user = make_ops<UserHam>();
static const HamOps* ham_ops(int ham_id) { return ham_id == JB200_HAM_USER ? &jb_ops_user : nullptr; }
#elif defined(JB_FAST_BUILD)
static const HamOps jb_ops_vdp = make_ops<HamVdp>();
static const HamOps* ham_ops(int ham_id) { return ham_id == JB200_HAM_VDP ? &jb_ops_vdp : nullptr; }
#else
extern const HamOps jb_ops_harmonic, jb_ops_vdp, jb_ops_vdp_pmp_smooth, jb_ops_vdp_pmp_bang, jb_ops_vdp_aug_bang, jb_ops_mlp;
static const HamOps* ham_ops(int ham_id) {
  switch (ham_id) {
    case JB200_HAM_HARMONIC: return &jb_ops_harmonic;
    case JB200_HAM_VDP: return &jb_ops_vdp;
    case JB200_HAM_VDP_PMP_SMOOTH: return &jb_ops_vdp_pmp_smooth;
    case JB200_HAM_VDP_PMP_BANG: return &jb_ops_vdp_pmp_bang;
    case JB200_HAM_VDP_AUG_BANG: return &jb_ops_vdp_aug_bang;
    case JB200_HAM_MLP: return &jb_ops_mlp;
    default: return nullptr;
  }
}
#endif

int lu16_solve_launch(long long M, const double* Are, const double* Aim, const double* bre, const double* bim, double* xre,
                      double* xim, cudaStream_t st);   // tu_mlp.cu

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
  P.Gustafsson = o->Gustafsson ? (1 | ((o->radaute_quirks & JB200_Q8_GUSTAFSSON_UNGATED) ? 2 : 0)) : 0;
  for (int i = 0; i < nhp && i < JB200_MAX_HAM_PARAMS; i++) P.ham_par[i] = hp[i];
}


static int solve_dev(Handle* H, const double* y0, const double* tspan, long long M, double* y_out, double* t_out,
                     double* h_out, int* res, long long* stats, int* nstg, long long* work, const DenseArgs& dense,
                     cudaStream_t st) {
  if (M <= 0) {
    H->launches = 0;
    return 0;
  }
  const HamOps* ops = ham_ops(H->ham_id);
  if (!ops) return fail(JB200_ENOTSUP, "unknown Hamiltonian id");
  return ops->solve(H, y0, tspan, M, y_out, t_out, h_out, res, stats, nstg, work, dense, st);
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

static __global__ void narrow_stats_kernel(const long long* __restrict__ s64, long long n, int* __restrict__ s32) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s32[i] = (int)s64[i];
}

static void free_workspace(Handle* H) {
  cudaFree(H->d_y0); cudaFree(H->d_ts); cudaFree(H->d_y); cudaFree(H->d_t); cudaFree(H->d_h);
  cudaFree(H->d_res); cudaFree(H->d_nstg); cudaFree(H->d_stats); cudaFree(H->d_work); cudaFree(H->d_stats32);
  H->d_y0 = H->d_ts = H->d_y = H->d_t = H->d_h = nullptr;
  H->d_res = H->d_nstg = H->d_stats32 = nullptr;
  H->d_stats = H->d_work = nullptr;
  H->cap = 0;
}

static int ensure_workspace(Handle* H, long long M) {
  if (M <= H->cap) return 0;
  free_workspace(H);   // pointers are nulled: a failed allocation below leaves nothing dangling
  const int ny = H->ny;
  struct Guard { Handle* H; bool ok = false; ~Guard() { if (!ok) { free_workspace(H); cudaGetLastError(); } } } guard{H};
  JB_CUDA(cudaMalloc(&H->d_stats32, sizeof(int) * M * 8));
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
  guard.ok = true;
  return 0;
}


}  // namespace jb

using namespace jb;

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
  o->radaute_quirks = JB200_Q8_GUSTAFSSON_UNGATED;
  return 0;
}

int jb200_create(const jb200_opts* o, int ham_id, const double* ham_params, int n_ham_params, void** handle) {
  if (!o || !handle) return fail(JB200_EINVAL, "null argument");
  const HamOps* ops = ham_ops(ham_id);
  if (!ops) return fail(JB200_ENOTSUP, "unknown Hamiltonian id (not compiled into this library)");
  const int ny = ops->ny;
  if (!ops->create && ops->npar > JB200_MAX_HAM_PARAMS)   // parameters live in RadauParams::ham_par[8] (device-resident ones: create)
    return fail(JB200_EINVAL, "this Hamiltonian declares more than JB200_MAX_HAM_PARAMS parameters");
  if (n_ham_params < ops->npar || (n_ham_params > 0 && !ham_params))
    return fail(JB200_EINVAL, "Hamiltonian needs " + std::to_string(ops->npar) + " parameters");
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
  if (ops->create) {
    const int rc2 = ops->create(H, ham_params, n_ham_params);
    if (rc2) {
      jb200_destroy(H);
      return rc2;
    }
  }
  if (const char* ep = std::getenv("JB200_EPOCH_PASSES")) H->epoch_passes = std::max(1, atoi(ep));
  if (const char* fe = std::getenv("JB200_FIXED_EPOCHS")) H->fixed_epochs = atoi(fe) != 0;
  if (const char* ps = std::getenv("JB200_PRESORT")) H->presort = atoi(ps) != 0;
  if (const char* eh = std::getenv("JB200_EPOCH_HORDER")) H->epoch_horder = atoi(eh);
  H->P.key_fine = 0;   // per batch: EpochWS::key_fine (solver_impl.cuh params_for)
  *handle = H;
  return 0;
}

void jb200_destroy(void* handle) {
  Handle* H = (Handle*)handle;
  if (!H) return;
  if (const HamOps* ops = ham_ops(H->ham_id))
    if (ops->destroy) ops->destroy(H);
  cudaFree(H->d_counter);
  ws_free(H->ws);
  cudaFree(H->d_y0); cudaFree(H->d_ts); cudaFree(H->d_y); cudaFree(H->d_t); cudaFree(H->d_h);
  cudaFree(H->d_res); cudaFree(H->d_nstg); cudaFree(H->d_stats); cudaFree(H->d_work); cudaFree(H->d_stats32);
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

int jb200_set_events(void* handle, int K, const double* A, const double* b, const int32_t* stop, const int32_t* slope) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (K < 0 || K > 4) return fail(JB200_EINVAL, "0 <= K <= 4 event functions");
  if (K > 0 && (!A || !b || !stop || !slope)) return fail(JB200_EINVAL, "null argument");
  if (H->ham_id == JB200_HAM_MLP && K > 0) return fail(JB200_ENOTSUP, "events are available for the thread-resident Hamiltonians only");
  H->P.ev_K = K;
  for (int k = 0; k < K; k++) {
    for (int i = 0; i < H->ny; i++) H->P.ev_A[k][i] = A[k * H->ny + i];
    H->P.ev_b[k] = b[k];
    H->P.ev_stop[k] = stop[k] != 0;
    H->P.ev_slope[k] = slope[k] > 0 ? 1 : (slope[k] < 0 ? -1 : 0);
  }
  return 0;
}

int jb200_solve_out(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out, double* t_out, double* h_out,
                    int32_t* res, int64_t* stats, int32_t* nstg, int64_t* work, int refine, int cap, double* tout, double* yout,
                    int32_t* nout, int ecap, double* te, double* ye, int32_t* ie, int32_t* nev, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M < 0) return fail(JB200_EINVAL, "M < 0");
  if (refine < 1 || cap < 0 || ecap < 0) return fail(JB200_EINVAL, "refine >= 1, cap >= 0, ecap >= 0");
  if (M > 0 && (!y0 || !tspan || !y_out || !t_out || !h_out || !res || !stats || !nout)) return fail(JB200_EINVAL, "null buffer");
  if (cap > 0 && (!tout || !yout)) return fail(JB200_EINVAL, "null output buffer");
  if (H->P.ev_K > 0 && (!nev || (ecap > 0 && (!te || !ye || !ie)))) return fail(JB200_EINVAL, "null event buffer");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  JB_CUDA(cudaMemsetAsync(nout, 0, sizeof(int) * M, st));
  if (nev) JB_CUDA(cudaMemsetAsync(nev, 0, sizeof(int) * M, st));
  DenseArgs d{2, cap > 0 ? tout : nullptr, cap > 0 ? yout : nullptr, (int*)nout};
  d.mode = refine > 1 ? 2 : 1;
  d.cap = cap;
  d.refine = refine;
  if (H->P.ev_K > 0) {
    d.te = te; d.ye = ye; d.ie = (int*)ie; d.nev = (int*)nev; d.ecap = ecap;
  }
  return solve_dev(H, y0, tspan, (long long)M, y_out, t_out, h_out, (int*)res, (long long*)stats, (int*)nstg,
                   (long long*)work, d, st);
}

static bool tspan_strictly_monotonic(const double* tspan, int64_t M) {
  bool ok = true;
  for (int64_t m = 0; m < M; m++) ok = ok && (tspan[2 * m + 1] != tspan[2 * m]);
  return ok;
}
// the check on a second host thread while the GPU solves; if no thread can be started it runs here and now
static std::future<bool> check_tspan_async(const double* tspan, int64_t M) {
  try {
    return std::async(std::launch::async, tspan_strictly_monotonic, tspan, M);
  } catch (...) {
    std::promise<bool> p;
    p.set_value(tspan_strictly_monotonic(tspan, M));
    return p.get_future();
  }
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
  int rc = ensure_workspace(H, M);
  if (rc) return rc;
  cudaStream_t st = H->own_stream;
  const int ny = H->ny;
  JB_CUDA(cudaMemcpyAsync(H->d_y0, y0, sizeof(double) * M * ny, cudaMemcpyHostToDevice, st));
  JB_CUDA(cudaMemcpyAsync(H->d_ts, tspan, sizeof(double) * M * 2, cudaMemcpyHostToDevice, st));
  // tspan must be strictly monotonic (radaute_impl.hpp:54-59): checked on the host copy while the GPU works (the pass over
  // 16 bytes per row costs ~1.5 ms per million rows on one core -- 2 % of the solve if it ran in front of it)
  auto span_ok = check_tspan_async(tspan, M);
  const DenseArgs none{2, nullptr, nullptr, nullptr};
  rc = solve_dev(H, H->d_y0, H->d_ts, M, H->d_y, H->d_t, H->d_h, H->d_res, H->d_stats, H->d_nstg, H->d_work, none, st);
  if (!span_ok.get()) {
    cudaStreamSynchronize(st);
    return fail(JB200_EINVAL, "Time vector must be strictly monotonic");
  }
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

int jb200_solve_host_compact(void* handle, const double* y0, const double* tspan, int64_t M, double* y_out, double* t_out,
                             double* h_out, int32_t* res, int32_t* stats32, int32_t* nstg, int64_t* work) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M < 0) return fail(JB200_EINVAL, "M < 0");
  if (M == 0) {
    H->launches = 0;
    return 0;
  }
  if (!y0 || !tspan || !y_out || !res || !stats32) return fail(JB200_EINVAL, "null buffer");
  int rc = ensure_workspace(H, M);
  if (rc) return rc;
  cudaStream_t st = H->own_stream;
  const int ny = H->ny;
  JB_CUDA(cudaMemcpyAsync(H->d_y0, y0, sizeof(double) * M * ny, cudaMemcpyHostToDevice, st));
  JB_CUDA(cudaMemcpyAsync(H->d_ts, tspan, sizeof(double) * M * 2, cudaMemcpyHostToDevice, st));
  auto span_ok = check_tspan_async(tspan, M);   // overlaps the solve (see jb200_solve_host)
  const DenseArgs none{2, nullptr, nullptr, nullptr};
  rc = solve_dev(H, H->d_y0, H->d_ts, M, H->d_y, H->d_t, H->d_h, H->d_res, H->d_stats, H->d_nstg, work ? H->d_work : nullptr,
                 none, st);
  if (!span_ok.get()) {
    cudaStreamSynchronize(st);
    return fail(JB200_EINVAL, "Time vector must be strictly monotonic");
  }
  if (rc) return rc;
  JB_CUDA(cudaMemcpyAsync(y_out, H->d_y, sizeof(double) * M * ny, cudaMemcpyDeviceToHost, st));
  narrow_stats_kernel<<<(unsigned)std::min<long long>((M * 8 + 255) / 256, 148 * 16), 256, 0, st>>>(H->d_stats, M * 8, H->d_stats32);
  JB_CUDA(cudaGetLastError());
  H->launches += 1;
  JB_CUDA(cudaMemcpyAsync(res, H->d_res, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaMemcpyAsync(stats32, H->d_stats32, sizeof(int) * M * 8, cudaMemcpyDeviceToHost, st));
  if (t_out) JB_CUDA(cudaMemcpyAsync(t_out, H->d_t, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
  if (h_out) JB_CUDA(cudaMemcpyAsync(h_out, H->d_h, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
  if (nstg) JB_CUDA(cudaMemcpyAsync(nstg, H->d_nstg, sizeof(int) * M, cudaMemcpyDeviceToHost, st));
  if (work) JB_CUDA(cudaMemcpyAsync(work, H->d_work, sizeof(long long) * M, cudaMemcpyDeviceToHost, st));
  JB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- incremental stepping
struct State {
  Handle* H;
  EpochWS ws;
};

int jb200_state_create(void* handle, const double* y0, const double* tspan, int64_t M, void** state, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H || !state) return fail(JB200_EINVAL, "null argument");
  if (M < 0 || (M > 0 && (!y0 || !tspan))) return fail(JB200_EINVAL, "bad batch");
  const HamOps* ops = ham_ops(H->ham_id);
  State* S = new State{H, EpochWS()};
  if (M > 0) {
    int rc = ops->st_init(H, S->ws, y0, tspan, 2, (long long)M, (cudaStream_t)stream);
    if (rc) {
      if (ops->st_free) ops->st_free(S->ws);
      else ws_free(S->ws);
      delete S;
      return rc;
    }
  }
  *state = S;
  return 0;
}

int jb200_state_step(void* state, const double* htry, void* stream) {
  State* S = (State*)state;
  if (!S) return fail(JB200_EINVAL, "null state");
  return ham_ops(S->H->ham_id)->st_step(S->H, S->ws, htry, (cudaStream_t)stream);
}

int jb200_state_read(void* state, double* y, double* t, double* hnext, double* hdid, int32_t* res, int64_t* stats,
                     int32_t* nstg, int32_t* done, void* stream) {
  State* S = (State*)state;
  if (!S) return fail(JB200_EINVAL, "null state");
  if (S->ws.M > 0 && (!y || !t || !hnext || !hdid || !res || !stats)) return fail(JB200_EINVAL, "null buffer");
  return ham_ops(S->H->ham_id)->st_read(S->H, S->ws, y, t, hnext, hdid, (int*)res, (long long*)stats, (int*)nstg,
                                        (int*)done, (cudaStream_t)stream);
}

void jb200_state_destroy(void* state) {
  State* S = (State*)state;
  if (!S) return;
  const HamOps* ops = ham_ops(S->H->ham_id);
  if (ops && ops->st_free) ops->st_free(S->ws);
  else ws_free(S->ws);
  delete S;
}

int jb200_last_launches(void* handle) { return handle ? ((Handle*)handle)->launches : 0; }

int jb200_eval_dyns(void* handle, const double* y, int64_t M, double* ydot, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  return ham_ops(H->ham_id)->dyns(H, y, M, ydot, (cudaStream_t)stream);
}

int jb200_eval_jac(void* handle, const double* y, int64_t M, double* jac, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  return ham_ops(H->ham_id)->jac(H, y, M, jac, (cudaStream_t)stream);
}

int jb200_solve_sens_par(void* handle, const double* y0, const double* tspan, const double* S0, int32_t n_par,
                         const int32_t* par_index, int64_t M, double* y_out, double* S_out, double* Sp_out, double* t_out,
                         double* h_out, int32_t* res, int64_t* stats, int32_t* nstg, int64_t* work, int32_t* sens_iters,
                         void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M < 0) return fail(JB200_EINVAL, "M < 0");
  if (M == 0) return 0;
  if (!y0 || !tspan || !y_out || !S_out || !t_out || !h_out || !res || !stats) return fail(JB200_EINVAL, "null buffer");
  const HamOps* ops = ham_ops(H->ham_id);
  if (!ops || !ops->sens) return fail(JB200_ENOTSUP, "forward sensitivities are available for the thread-resident Hamiltonians only");
  if (n_par < 0 || n_par > JB_SENS_MAXPAR) return fail(JB200_EINVAL, "at most 4 parameter directions per solve");
  if (n_par > 0 && (!par_index || !Sp_out)) return fail(JB200_EINVAL, "parameter directions need par_index and Sp_out");
  for (int k = 0; k < n_par; k++)
    if (par_index[k] < 0 || par_index[k] >= ops->npar) return fail(JB200_EINVAL, "parameter index out of range for this Hamiltonian");
  H->P.sens_npar = n_par;
  for (int k = 0; k < JB_SENS_MAXPAR; k++) H->P.sens_par[k] = k < n_par ? par_index[k] : 0;
  const int rc = ops->sens(H, y0, tspan, S0, (long long)M, y_out, S_out, t_out, h_out, (int*)res, (long long*)stats, (int*)nstg,
                           (long long*)work, (int*)sens_iters, Sp_out, (cudaStream_t)stream);
  H->P.sens_npar = 0;
  return rc;
}

int jb200_solve_sens(void* handle, const double* y0, const double* tspan, const double* S0, int64_t M, double* y_out,
                     double* S_out, double* t_out, double* h_out, int32_t* res, int64_t* stats, int32_t* nstg,
                     int64_t* work, int32_t* sens_iters, void* stream) {
  return jb200_solve_sens_par(handle, y0, tspan, S0, 0, nullptr, M, y_out, S_out, nullptr, t_out, h_out, res, stats, nstg, work,
                              sens_iters, stream);
}

int jb200_dopri853_step(void* handle, int64_t M, int64_t norm_M, double atol, double rtol, int aliased, double* x, double* y,
                        double* dydx, double* errold, double* hnext, const double* htry, double* hdid, int32_t* nrej,
                        void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  if (!x || !y || !dydx || !errold || !hnext || !htry || !hdid) return fail(JB200_EINVAL, "null buffer");
  if (!(atol >= 0) || !(rtol >= 0) || (atol == 0 && rtol == 0)) return fail(JB200_EINVAL, "atol, rtol must be >= 0 and not both 0");
  const HamOps* ops = ham_ops(H->ham_id);
  if (!ops || !ops->dopri) return fail(JB200_ENOTSUP, "DOPRI853 is available for the thread-resident Hamiltonians only");
  const dp::StepArgs a{(long long)M, (long long)(norm_M > 0 ? norm_M : M), atol, rtol, x, y, dydx, errold, hnext, aliased, htry, hdid,
                       (int*)nrej};
  return ops->dopri(H, a, (cudaStream_t)stream);
}

int jb200_dopri853_step_dual(void* handle, int64_t M, int64_t norm_M, double atol, double rtol, int aliased, double* x, double* y,
                             double* dydx, double* errold, double* hnext, const double* htry, double* hdid, int32_t* nrej,
                             double* yd, int32_t nd, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  if (!x || !y || !dydx || !errold || !hnext || !htry || !hdid || !yd) return fail(JB200_EINVAL, "null buffer");
  if (nd < 1 || nd > dp::NDMAX) return fail(JB200_EINVAL, "1 <= nd <= 8 dual directions per call");
  if (!(atol >= 0) || !(rtol >= 0) || (atol == 0 && rtol == 0)) return fail(JB200_EINVAL, "atol, rtol must be >= 0 and not both 0");
  const HamOps* ops = ham_ops(H->ham_id);
  if (!ops || !ops->dopri) return fail(JB200_ENOTSUP, "DOPRI853 is available for the thread-resident Hamiltonians only");
  dp::StepArgs a{(long long)M, (long long)(norm_M > 0 ? norm_M : M), atol, rtol, x, y, dydx, errold, hnext, aliased, htry, hdid,
                 (int*)nrej};
  a.yd = yd;
  a.nd = nd;
  return ops->dopri(H, a, (cudaStream_t)stream);
}

int jb200_eval_jac_dual(void* handle, const double* y, const double* yd, int ND, int64_t M, double* f, double* J, double* T,
                        double* fd, double* Jd, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  if (!y || !f || !J) return fail(JB200_EINVAL, "null buffer");
  if (ND < 0 || (ND > 0 && !yd)) return fail(JB200_EINVAL, "ND directions need yd");
  const HamOps* ops = ham_ops(H->ham_id);
  if (!ops || !ops->jac_dual) return fail(JB200_ENOTSUP, "third-order evaluation is available for the thread-resident Hamiltonians only");
  const JacDualArgs a{(long long)M, ND, y, ND > 0 ? yd : nullptr, f, J, T, ND > 0 ? fd : nullptr, ND > 0 ? Jd : nullptr};
  return ops->jac_dual(H, a, (cudaStream_t)stream);
}

int jb200_seulex_step(void* handle, int64_t M, double atol, double rtol, int KMAXX, int quirks, const double* htry, double* y,
                      double* x, double* theta, int32_t* k_targ, double* hnext, double* hdid, int32_t* flags, double* yjac,
                      int32_t* k_out, int32_t* nrej, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  if (!htry || !y || !x || !theta || !k_targ || !hnext || !hdid || !flags || !yjac) return fail(JB200_EINVAL, "null buffer");
  if (!(atol >= 0) || !(rtol >= 0) || (atol == 0 && rtol == 0)) return fail(JB200_EINVAL, "atol, rtol must be >= 0 and not both 0");
  if (KMAXX < 2 || KMAXX + 1 > sx::IMAXX_CAP) return fail(JB200_EINVAL, "KMAXX must be in [2, 15]");
  const HamOps* ops = ham_ops(H->ham_id);
  if (!ops || !ops->seulex) return fail(JB200_ENOTSUP, "SEULEX is available for the thread-resident Hamiltonians only");
  const sx::StepArgs a{(long long)M, atol, rtol, KMAXX, quirks, htry, y, x, theta, (int*)k_targ, hnext, hdid, (int*)flags, yjac,
                       (int*)k_out, (int*)nrej};
  return ops->seulex(H, a, (cudaStream_t)stream);
}

int jb200_seulex_step_dual(void* handle, int64_t M, double atol, double rtol, int KMAXX, int quirks, const double* htry, double* y,
                           double* x, double* theta, int32_t* k_targ, double* hnext, double* hdid, int32_t* flags, double* yjac,
                           int32_t* k_out, int32_t* nrej, double* yd, double* sjac, int32_t nd, void* stream) {
  Handle* H = (Handle*)handle;
  if (!H) return fail(JB200_EINVAL, "null handle");
  if (M <= 0) return 0;
  if (!htry || !y || !x || !theta || !k_targ || !hnext || !hdid || !flags || !yjac || !yd || !sjac) return fail(JB200_EINVAL, "null buffer");
  if (nd < 1 || nd > sx::NDMAX) return fail(JB200_EINVAL, "1 <= nd <= 8 dual directions per call");
  if (!(atol >= 0) || !(rtol >= 0) || (atol == 0 && rtol == 0)) return fail(JB200_EINVAL, "atol, rtol must be >= 0 and not both 0");
  if (KMAXX < 2 || KMAXX + 1 > sx::IMAXX_CAP) return fail(JB200_EINVAL, "KMAXX must be in [2, 15]");
  const HamOps* ops = ham_ops(H->ham_id);
  if (!ops || !ops->seulex) return fail(JB200_ENOTSUP, "SEULEX is available for the thread-resident Hamiltonians only");
  sx::StepArgs a{(long long)M, atol, rtol, KMAXX, quirks, htry, y, x, theta, (int*)k_targ, hnext, hdid, (int*)flags, yjac,
                 (int*)k_out, (int*)nrej};
  a.yd = yd;
  a.sjac = sjac;
  a.nd = nd;
  return ops->seulex(H, a, (cudaStream_t)stream);
}

int jb200_seulex_init(double atol, double rtol, int KMAXX, double* theta0, int32_t* k_targ0) {
  if (!theta0 || !k_targ0) return fail(JB200_EINVAL, "null argument");
  if (!(atol >= 0) || !(rtol >= 0) || (atol == 0 && rtol == 0)) return fail(JB200_EINVAL, "atol, rtol must be >= 0 and not both 0");
  // SeulexState.__init__ (src/python/seulext.py:44-46, 66-70): theta = 2 jac_redo, k_targ from the tolerances
  const double jac_redo = std::fmin(1.0e-4, rtol);
  *theta0 = 2.0 * jac_redo;
  const double logfact = -std::log10(rtol + atol) * 0.6 + 0.5;
  *k_targ0 = (int)std::floor(std::fmax(1.0, std::fmin((double)(KMAXX - 1), std::nearbyint(logfact))));
  return 0;
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
    case 5: return launch_lu<5>(M, Are, Aim, bre, bim, xre, xim, st);
    case 6: return launch_lu<6>(M, Are, Aim, bre, bim, xre, xim, st);
    case 8: return launch_lu<8>(M, Are, Aim, bre, bim, xre, xim, st);   // Ny = 8: user Hamiltonians with N = 4
#if !defined(JB_USER_HAM_HEADER) && !defined(JB_FAST_BUILD)
    case 16: return lu16_solve_launch(M, Are, Aim, bre, bim, xre, xim, st);   // 16 lanes per system (csrc/lu16.cuh)
#endif
    default: return fail(JB200_ENOTSUP, "jb200_lu_solve: n must be 2, 3, 4, 5, 6, 8 or 16");
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

int jb200_dmma_peak(int iters, double* tflops) {
  if (!tflops || iters < 1) return fail(JB200_EINVAL, "bad argument");
  int dev = 0, sms = 148;
  JB_CUDA(cudaGetDevice(&dev));
  JB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  JB_CUDA(cudaMalloc(&sink, sizeof(double)));
  cudaEvent_t e0, e1;
  JB_CUDA(cudaEventCreate(&e0));
  JB_CUDA(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  dmma_kernel<<<blocks, threads>>>(iters, 1.0, sink);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    JB_CUDA(cudaEventRecord(e0));
    dmma_kernel<<<blocks, threads>>>(iters, 1.0, sink);
    JB_CUDA(cudaEventRecord(e1));
    JB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    JB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 512.0 * 8.0 * (double)iters * (double)blocks * (double)(threads / 32);   // 2*8*8*4 per warp-level mma
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
