// Stage evaluation for an MLP-parameterised Hamiltonian (config 4 of BASELINE.json):
//   H(v) = W4 . tanh(W3 tanh(W2 tanh(W1 v + b1) + b2) + b3) + b4,   v = [x ; p] (16 inputs),
//   width 128, fp64.
// Replaces, for this Hamiltonian, the reference's autograd ladders
//   evalDyns  /root/reference/src/cpp/janus_ode_common.hpp:2161-2186  (3 forwards + 2 backwards of H)
//   evalJac   /root/reference/src/cpp/janus_ode_common.hpp:2220-2256  (5 forwards + 4N row backwards)
// by two tile kernels in which the batch dimension makes every layer a dense GEMM:
//   mlp_grad_kernel  64 (trajectory, stage) rows per CTA: forward (3 GEMMs + tanh), reverse sweep
//                    (2 transposed GEMMs + the 128x16 output product) -> ydot = [sg dH/dx ; dH/dp]
//   mlp_hess_kernel  4 trajectories x 16 tangent directions per CTA: forward-over-reverse (tangent
//                    forward 2 GEMMs, tangent reverse 2 GEMMs + output product) -> the 16x16 Jacobian
// One CTA = 256 threads = two warps per SM sub-partition.  The layer products run on the FP64 TENSOR CORES
// (mma.sync.m8n8k4.f64, "DMMA"): a warp owns a 32-feature x 32-row block of the 64 x 128 layer output = 4 x 4 mma
// tiles, 32 accumulators per thread; per k-step of 4 it loads 4 weight fragments + 4 activation fragments (one double
// each, bank-conflict-free through an XOR swizzle of both shared-memory operands) for 16 DMMAs = 128 DFMA-equivalents --
// 6x fewer shared-memory instructions and 8x fewer issue slots per multiply-add than the register-tiled DFMA GEMM it
// replaces (54-61 % FP64-pipe active).  The activations of the tile stay in shared memory ([feature][row], 3 x 64 KB)
// from the first layer to the last, weights stream from L2 in 16-row chunks through a cp.async double buffer.
//
// ARITHMETIC CONTRACT (mirrored by oracle/c/radau_oracle.c mlp_eval_mirror): every dot product is
// ONE fma chain in increasing index order, seeded with the bias (forward) or 0.0 (transposed);
// tanh = det_tanh; 1 - a^2 = fma(-a, a, 1).  DMMA keeps that contract: D = A B + C is, per output, the chain
// fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0, c)))) -- measured on B200 against scalar fma on 262,144 outputs with a
// 2^40 dynamic range, all bit-identical (profiles/probes/dmma_order.cu) -- and successive k-steps chain through the
// accumulator, so results stay bit-identical to the scalar mirror.
#pragma once
#include <cuda_runtime.h>

#include "detmath.cuh"
#include "mlp_common.cuh"

namespace jb {
namespace mlp {

// Tile shape (JB_MLP_CFG):
//   2            64 rows per CTA, 512 threads = 16 warps (4 per SM sub-partition), 224 KB, one CTA per SM: the GEMM loops
//                saturate the FP64 tensor pipe with 2 warps per sub-partition already; the phases between them (tanh sweeps =
//                dependent fma chains, epilogues, output product, chunk barriers) need the extra warps to hide latency
//   1            32 rows per CTA, 256 threads, 112 KB, TWO CTAs per SM (one CTA's pipe-light phase overlaps the other's GEMM):
//                twice the L2 -> shared-memory weight traffic per row (every CTA streams all weights per tile)
//   0 (default)  64 rows, 256 threads, one CTA per SM.  Measured on config 4 (2^20 rows, with the branch-free tanh): 0 / 1 / 2 =
//                1008.8 / 1009.1 / 1023.6 ms per solve -- the tile kernels are no longer what separates the shapes
#ifndef JB_MLP_CFG
#define JB_MLP_CFG 0
#endif
constexpr int TR = JB_MLP_CFG == 1 ? 32 : 64;          // rows per CTA tile
constexpr int KC = JB_MLP_CFG == 1 ? 8 : 16;           // weight rows per cp.async chunk
constexpr int CTAS_PER_SM = JB_MLP_CFG == 1 ? 2 : 1;
constexpr int THREADS = JB_MLP_CFG == 2 ? 512 : 256;   // warps = 4 feature blocks (32 features each) x WR row blocks
constexpr int WR = THREADS / 32 / 4;
constexpr int ACT_WORDS = WD * TR;                                      // one activation buffer
constexpr size_t SMEM_BYTES = sizeof(double) * (3 * ACT_WORDS + 2 * KC * WD);   // 112 KB (TR = 32) / 224 KB (TR = 64)

// device-resident weights of one Hamiltonian (both orientations of every matrix, k-major for the
// GEMM that consumes it)
struct Weights {
  const double* W1t;   // [16][128]   W1t[k][c] = W1[c][k]      forward layer 1
  const double* W2t;   // [128][128]  W2t[k][c] = W2[c][k]      forward layer 2
  const double* W3t;   // [128][128]
  const double* W3;    // [128][128]  natural (out-major): reverse through layer 3
  const double* W2;    // [128][128]
  const double* W1;    // [128][16]   natural: output product
  const double* b1;    // [128]
  const double* b2;
  const double* b3;
  const double* W4;    // [128]
};

// Shared-memory operand layout: element (k, i) of a [K][LD] operand (activations: LD = TR rows; weights: LD = WD
// features) lives at k * LD + (i ^ x(k)), x(k) = 8 (k & 1) + 4 ((k >> 1) & 1).  A fragment load touches, per half-warp, 4
// values of k x 4 consecutive i: without the swizzle those 16 doubles share 8 banks (LD * 8 bytes is a multiple of 128), with
// it they are 16 distinct consecutive doubles = one conflict-free 128-byte wavefront.  Pairs (i even, i + 1) stay adjacent,
// and putting the LOW bit of k on the 64-byte bit makes the epilogue stores conflict-free too: a quarter-warp stores the row
// pairs of two neighbouring features, which land in different halves of the 128-byte bank row (with x(k) = 4 (k & 3) they
// shared one: 3.7- / 7.6-way store conflicts in the first DMMA version).
__device__ __forceinline__ int swz_x(int k) { return ((k & 1) << 3) | ((k & 2) << 1); }
__device__ __forceinline__ int swz(int k, int i) { return i ^ swz_x(k); }
__device__ __forceinline__ int aidx(int c, int r) { return c * TR + swz(c, r); }   // activation buffers [feature][row]

// thread -> accumulators: warp w owns rows (TR / WR) (w % WR) .., features 32 (w / WR) ..; mma tile (j, f), j < NJ, f < NF,
// covers rows rb + 8 j .. + 7 and features cb + 8 f .. + 7; this lane holds, of each tile, feature col(f) and the row
// pair rowp(j), +1
struct TileMap {
  int rb, cb, lq, lr;   // lq = lane >> 2 (fragment row / column index), lr = lane & 3 (k index within a k-step)
  __device__ TileMap() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    rb = (TR / WR) * (warp % WR);
    cb = 32 * (warp / WR);
    lq = lane >> 2;
    lr = lane & 3;
  }
  __device__ __forceinline__ int rowp(int j) const { return rb + 8 * j + 2 * lr; }   // even row of this lane's pair, j < NJ
  __device__ __forceinline__ int col(int f) const { return cb + 8 * f + lq; }        // f < 4
};
constexpr int NJ = TR / WR / 8, NF = 4;   // mma tiles per warp: row tiles x feature tiles
constexpr int NSL = (NJ + 1) / 2;      // Hessian kernel: trajectory slots (16 tangent rows each) a thread's rows belong to

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D (8 features x 8 rows) += A (8 features x 4 k: the weights) * B (4 k x 8 rows: the activations), FP64 tensor cores
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// acc[j][f][e] += sum_k A[k][rowp(j) + e] * B[k][col(f)], k = 0..K-1 in order (one fma chain each).
// A: shared [K][TR], swizzled; B: global [K][WD] (row-major, 16-byte aligned); wbuf: shared [2][KC][WD], swizzled.
__device__ __forceinline__ void gemm_tile(const double* __restrict__ A, const double* __restrict__ B, int K,
                                          double* __restrict__ wbuf, const TileMap& tm, double (&acc)[NJ][NF][2]) {
  const int nchunks = K / KC;
  auto issue = [&](int ch) {
    const double* src = B + (size_t)ch * KC * WD;
    double* dst = wbuf + (ch & 1) * KC * WD;
#pragma unroll
    for (int t = 0; t < (KC * WD / 2) / THREADS; t++) {
      const int o = (t * THREADS + threadIdx.x) * 2;
      const int kk = o / WD, c = o % WD;
      cp_async16(dst + kk * WD + swz(kk, c), src + o);
    }
    cp_async_commit();
  };
  issue(0);
  // this lane's fragment offsets within a k-step (k = lr): weights (k, cb + 8 f + lq), activations (k, rb + 8 j + lq)
  const int xs = swz_x(tm.lr);
  for (int ch = 0; ch < nchunks; ch++) {
    if (ch + 1 < nchunks) {
      issue(ch + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* wb = wbuf + (ch & 1) * KC * WD + tm.lr * WD;
    const double* Ak = A + (size_t)ch * KC * TR + tm.lr * TR;
#pragma unroll
    for (int ks = 0; ks < KC / 4; ks++) {   // KC is a multiple of 4, so (k & 3) == lr for every k-step
      double a[NF], b[NJ];
#pragma unroll
      for (int f = 0; f < NF; f++) a[f] = wb[ks * 4 * WD + ((tm.cb + 8 * f + tm.lq) ^ xs)];
#pragma unroll
      for (int j = 0; j < NJ; j++) b[j] = Ak[ks * 4 * TR + ((tm.rb + 8 * j + tm.lq) ^ xs)];
#pragma unroll
      for (int j = 0; j < NJ; j++)
#pragma unroll
        for (int f = 0; f < NF; f++) dmma(acc[j][f], a[f], b[j]);
    }
    __syncthreads();   // everyone is done with this chunk before it is overwritten
  }
}

// out[row][j] = sum_i W1[i][j] * D[i][row], j < 16, on the tensor cores as well: 2 feature tiles x TR / 8 row tiles, OPT of
// them per warp (the same row tile, so the activation fragment is loaded once); the weight fragments W1[i][j] (natural
// [128][16], 16 KB, the same for every CTA) come straight from L1 / L2.  Chain order: i increasing, seeded with 0.
constexpr int ORT = TR / 8;                          // row tiles
constexpr int OPT = (2 * ORT) / (THREADS / 32);      // tiles per warp: 2 (same row tile, both feature tiles) or 1
static_assert(OPT == 1 || OPT == 2, "output product tiling");
struct OutTile {
  int rt, jt0;
  __device__ OutTile() {
    const int warp = threadIdx.x >> 5;
    rt = warp % ORT;
    jt0 = OPT == 2 ? 0 : warp / ORT;
  }
  // this lane's results: feature(t) = 8 (jt0 + t) + lq, rows row0 + e, e < 2
  __device__ __forceinline__ int feature(int t, const TileMap& tm) const { return 8 * (jt0 + t) + tm.lq; }
  __device__ __forceinline__ int row0(const TileMap& tm) const { return 8 * rt + 2 * tm.lr; }
};
__device__ __forceinline__ void output_product(const double* __restrict__ D, const double* __restrict__ W1g, const TileMap& tm,
                                               const OutTile& ot, double (&g)[OPT][2]) {
#pragma unroll
  for (int t = 0; t < OPT; t++) g[t][0] = g[t][1] = 0.0;
  const int xs = swz_x(tm.lr);
  const double* Dk = D + tm.lr * TR;
  const double* Wk = W1g + tm.lr * NV + tm.lq;
#pragma unroll 8
  for (int ks = 0; ks < WD / 4; ks++) {
    const double b = Dk[ks * 4 * TR + ((8 * ot.rt + tm.lq) ^ xs)];
#pragma unroll
    for (int t = 0; t < OPT; t++) dmma(g[t], __ldg(Wk + ks * 4 * NV + 8 * (ot.jt0 + t)), b);
  }
}

// store the accumulators raw into a [feature][row] buffer
__device__ __forceinline__ void store_tile(double* __restrict__ A, const TileMap& tm, const double (&acc)[NJ][NF][2]) {
#pragma unroll
  for (int f = 0; f < NF; f++)
#pragma unroll
    for (int j = 0; j < NJ; j++)
      *reinterpret_cast<double2*>(A + aidx(tm.col(f), tm.rowp(j))) = make_double2(acc[j][f][0], acc[j][f][1]);
}

template <class F>
__device__ __forceinline__ void seed_acc(double (&acc)[NJ][NF][2], const TileMap& tm, F bias) {
#pragma unroll
  for (int f = 0; f < NF; f++) {
    const double b = bias(tm.col(f));
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[j][f][0] = acc[j][f][1] = b;
  }
}

// a = tanh(a) over a whole buffer: a ROLLED loop (one copy of det_tanh_t in the instruction stream
// instead of 32 per layer), TI independent elements in flight per thread -- det_tanh_t is one long
// dependent FMA chain, the FP64 pipe needs ~16 of them per sub-partition to stay busy
constexpr int TI = 8;
__device__ __forceinline__ void tanh_buffer(double* __restrict__ A, const double* __restrict__ tab) {
#pragma unroll 1
  for (int o = threadIdx.x; o < ACT_WORDS; o += TI * THREADS) {
    double t[TI];
#pragma unroll
    for (int u = 0; u < TI; u++) t[u] = A[o + u * THREADS];
#pragma unroll
    for (int u = 0; u < TI; u++) t[u] = det_tanh_t(t[u], tab);
#pragma unroll
    for (int u = 0; u < TI; u++) A[o + u * THREADS] = t[u];
  }
}

// Where a row of the gradient kernel comes from / goes to.  Element e of trajectory m of an array
// lives at base[m * sr + e * se] (SoA state: sr = 1, se = M;  row-major API buffers: sr = 16, se = 1).
struct GradIO {
  const int* rows;         // [n_rows] row descriptors: trajectory * 8 + stage  (nullptr: row r = trajectory r, stage 0)
  const int* n_rows_dev;   // device count (nullptr: use n_rows)
  int n_rows;
  const double* y;         // state
  long long y_sr, y_se;
  const double* add;       // optional addend: add[(stage * 16 + e) * add_se + m * add_sr]  (stage increments z_q)
  long long add_sr, add_se;
  const double* add_scalar;  // optional scalar added to every component ([m]; Estrad's y + err)
  double* out;             // out[(stage * 16 + e) * out_se + m * out_sr]
  long long out_sr, out_se;
  double* act;             // optional [m][5][128] saved activations a1, a2, a3, q1, q2 (for the Jacobian)
  double sign;             // costate sign: ydot = [sign dH/dx ; dH/dp]
};

__global__ void __launch_bounds__(THREADS, CTAS_PER_SM) mlp_grad_kernel(const Weights Wt, const GradIO io) {
  extern __shared__ __align__(16) double smem[];
  double* A0 = smem;
  double* A1 = A0 + ACT_WORDS;
  double* A2 = A1 + ACT_WORDS;
  double* wbuf = A2 + ACT_WORDS;
  __shared__ int s_traj[TR], s_stage[TR];
  __shared__ double s_tab[32];   // 2^(j/32) for det_tanh_t: a run-time-indexed table belongs in shared memory
  const int n_rows = io.n_rows_dev ? *io.n_rows_dev : io.n_rows;
  const TileMap tm;
  if (threadIdx.x < 32) s_tab[threadIdx.x] = det_exp2_tab()[threadIdx.x];
  for (int tile = blockIdx.x; (long long)tile * TR < n_rows; tile += gridDim.x) {
    const int base = tile * TR;
    if (threadIdx.x < TR) {
      const int r = base + threadIdx.x;
      int d = -1;
      if (r < n_rows) d = io.rows ? io.rows[r] : r * 8;
      s_traj[threadIdx.x] = d < 0 ? -1 : d >> 3;
      s_stage[threadIdx.x] = d < 0 ? 0 : d & 7;
    }
    __syncthreads();
    // v = [x ; p] of every row -> A2[k][row] (k < 16); state order is y = [p ; x]
    {
      double v[NV * TR / THREADS], ad[NV * TR / THREADS];
#pragma unroll
      for (int u = 0; u < NV * TR / THREADS; u++) {   // all loads in flight before the first use
        const int o = threadIdx.x + u * THREADS;
        const int r = o & (TR - 1), e = ((o / TR) + NH) & (NV - 1);
        const int m = s_traj[r];
        v[u] = 0.0;
        ad[u] = 0.0;
        if (m >= 0) {
          v[u] = io.y[(size_t)m * io.y_sr + (size_t)e * io.y_se];
          if (io.add) ad[u] = io.add[(size_t)(s_stage[r] * NV + e) * io.add_se + (size_t)m * io.add_sr];
          else if (io.add_scalar) ad[u] = io.add_scalar[m];
        }
      }
#pragma unroll
      for (int u = 0; u < NV * TR / THREADS; u++) {
        const int o = threadIdx.x + u * THREADS;
        A2[aidx(o / TR, o & (TR - 1))] = (io.add || io.add_scalar) ? v[u] + ad[u] : v[u];
      }
    }
    __syncthreads();
    double acc[NJ][NF][2];
    // ---- layer 1: a1 = tanh(W1 v + b1) -> A0
    seed_acc(acc, tm, [&](int c) { return Wt.b1[c]; });
    gemm_tile(A2, Wt.W1t, NV, wbuf, tm, acc);
    store_tile(A0, tm, acc);
    __syncthreads();
    tanh_buffer(A0, s_tab);
    __syncthreads();
    // ---- layer 2 -> A1
    seed_acc(acc, tm, [&](int c) { return Wt.b2[c]; });
    gemm_tile(A0, Wt.W2t, WD, wbuf, tm, acc);
    store_tile(A1, tm, acc);
    __syncthreads();
    tanh_buffer(A1, s_tab);
    __syncthreads();
    // ---- layer 3: a3 = tanh(.), d3 = W4 (1 - a3^2) -> A2
    seed_acc(acc, tm, [&](int c) { return Wt.b3[c]; });
    gemm_tile(A1, Wt.W3t, WD, wbuf, tm, acc);
    store_tile(A2, tm, acc);
    __syncthreads();
    tanh_buffer(A2, s_tab);
    __syncthreads();
    for (int o = threadIdx.x; o < ACT_WORDS; o += THREADS) {
      const int c = o / TR, r = swz(c, o & (TR - 1));   // physical slot o holds logical (feature c, row r)
      const double a3 = A2[o];
      if (io.act && s_traj[r] >= 0) io.act[((size_t)s_traj[r] * 5 + 2) * WD + c] = a3;
      A2[o] = __ldg(Wt.W4 + c) * fma(-a3, a3, 1.0);
    }
    __syncthreads();
    // ---- reverse through layer 3: q2 = W3^T d3, d2 = q2 (1 - a2^2) -> A1 (in place)
    seed_acc(acc, tm, [](int) { return 0.0; });
    gemm_tile(A2, Wt.W3, WD, wbuf, tm, acc);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int c = tm.col(f);
#pragma unroll
      for (int j = 0; j < NJ; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int r = tm.rowp(j) + e;
          const double a2 = A1[aidx(c, r)], q2 = acc[j][f][e];
          if (io.act && s_traj[r] >= 0) {
            io.act[((size_t)s_traj[r] * 5 + 1) * WD + c] = a2;
            io.act[((size_t)s_traj[r] * 5 + 4) * WD + c] = q2;
          }
          A1[aidx(c, r)] = q2 * fma(-a2, a2, 1.0);
        }
    }
    __syncthreads();
    // ---- reverse through layer 2: q1 = W2^T d2, d1 = q1 (1 - a1^2) -> A0 (in place)
    seed_acc(acc, tm, [](int) { return 0.0; });
    gemm_tile(A1, Wt.W2, WD, wbuf, tm, acc);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int c = tm.col(f);
#pragma unroll
      for (int j = 0; j < NJ; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int r = tm.rowp(j) + e;
          const double a1 = A0[aidx(c, r)], q1 = acc[j][f][e];
          if (io.act && s_traj[r] >= 0) {
            io.act[((size_t)s_traj[r] * 5 + 0) * WD + c] = a1;
            io.act[((size_t)s_traj[r] * 5 + 3) * WD + c] = q1;
          }
          A0[aidx(c, r)] = q1 * fma(-a1, a1, 1.0);
        }
    }
    __syncthreads();
    // ---- g = W1^T d1 ; ydot = [sign dH/dx ; dH/dp]
    {
      const OutTile ot;
      double g[OPT][2];
      output_product(A0, Wt.W1, tm, ot, g);
#pragma unroll
      for (int e2 = 0; e2 < 2; e2++) {
        const int r = ot.row0(tm) + e2;
        const int m = s_traj[r];
        if (m >= 0) {
          const int st = s_stage[r];
#pragma unroll
          for (int t = 0; t < OPT; t++) {
            const int e = ot.feature(t, tm);
            const double v = (e < NH) ? io.sign * g[t][e2] : g[t][e2];
            io.out[(size_t)(st * NV + e) * io.out_se + (size_t)m * io.out_sr] = v;
          }
        }
      }
    }
    __syncthreads();
  }
}

// Jacobian J = d ydot / d y (rows [pdot ; xdot], columns [p ; x]) of the trajectories in `list`,
// from the activations saved by the gradient kernel at the same state.
struct HessIO {
  const int* list;         // [n] trajectory indices (nullptr: identity)
  const int* n_dev;        // device count (nullptr: use n)
  int n;
  const double* act;       // [m][5][128]: a1, a2, a3, q1, q2
  double* J;               // J[(r * 16 + c) * J_se + m * J_sr]
  long long J_sr, J_se;
  double sign;
};

// the epilogue operands of one layer (saved activations of this thread's 4 features x 2 trajectory slots), loaded
// BEFORE the layer's GEMM so that their L2 latency hides behind it
__device__ __forceinline__ void prefetch_act(const double* const (&actj)[NSL], int layer, const TileMap& tm,
                                             double (&v)[NSL][NF]) {
#pragma unroll
  for (int jj = 0; jj < NSL; jj++)
#pragma unroll
    for (int f = 0; f < NF; f++) v[jj][f] = __ldg(actj[jj] + layer * WD + tm.col(f));
}

__global__ void __launch_bounds__(THREADS, CTAS_PER_SM) mlp_hess_kernel(const Weights Wt, const HessIO io) {
  extern __shared__ __align__(16) double smem[];
  double* A0 = smem;
  double* A1 = A0 + ACT_WORDS;
  double* A2 = A1 + ACT_WORDS;
  double* wbuf = A2 + ACT_WORDS;
  constexpr int TPT = TR / 16;   // trajectories per tile (16 tangent rows each)
  __shared__ int s_traj[TPT];
  const int n = io.n_dev ? *io.n_dev : io.n;
  const TileMap tm;
  for (int tile = blockIdx.x; (long long)tile * TPT < n; tile += gridDim.x) {
    if (threadIdx.x < TPT) {
      const int q = tile * TPT + threadIdx.x;
      s_traj[threadIdx.x] = q < n ? (io.list ? io.list[q] : q) : -1;
    }
    __syncthreads();
    // rows of this thread: row r = rowp(j) + e = rb + 8 j + 2 lr + e  ->  trajectory slot r / 16 = (rb >> 4) + (j >> 1),
    // tangent direction k = r & 15 = 8 (j & 1) + 2 lr + e
    const double* actj[NSL];
#pragma unroll
    for (int jj = 0; jj < NSL; jj++) {
      const int sl = (tm.rb >> 4) + jj;
      actj[jj] = io.act + (size_t)(s_traj[sl] < 0 ? s_traj[0] : s_traj[sl]) * 5 * WD;
    }
    // ---- ta1[c][row] = (1 - a1^2) W1[c][k] -> A0
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int c = tm.col(f);
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const double a1 = actj[j >> 1][0 * WD + c];
        const double e1 = fma(-a1, a1, 1.0);
        const int k0 = 8 * (j & 1) + 2 * tm.lr;
        *reinterpret_cast<double2*>(A0 + aidx(c, tm.rowp(j))) =
            make_double2(e1 * Wt.W1[c * NV + k0], e1 * Wt.W1[c * NV + k0 + 1]);
      }
    }
    __syncthreads();
    double acc[NJ][NF][2];
    // ---- ta2 = (1 - a2^2) (W2 ta1) -> A1
    seed_acc(acc, tm, [](int) { return 0.0; });
    double pa[NSL][NF], pq[NSL][NF];
    prefetch_act(actj, 1, tm, pa);
    gemm_tile(A0, Wt.W2t, WD, wbuf, tm, acc);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int c = tm.col(f);
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const double a2 = pa[j >> 1][f];
        const double e2 = fma(-a2, a2, 1.0);
        *reinterpret_cast<double2*>(A1 + aidx(c, tm.rowp(j))) = make_double2(e2 * acc[j][f][0], e2 * acc[j][f][1]);
      }
    }
    __syncthreads();
    // ---- ta3 = (1 - a3^2) (W3 ta2);  td3 = W4 ((-2 a3) ta3) -> A2
    seed_acc(acc, tm, [](int) { return 0.0; });
    prefetch_act(actj, 2, tm, pa);
    gemm_tile(A1, Wt.W3t, WD, wbuf, tm, acc);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int c = tm.col(f);
      const double w4 = __ldg(Wt.W4 + c);
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const double a3 = pa[j >> 1][f];
        const double e3 = fma(-a3, a3, 1.0), m3 = -2.0 * a3;
        const double t0 = e3 * acc[j][f][0], t1 = e3 * acc[j][f][1];
        *reinterpret_cast<double2*>(A2 + aidx(c, tm.rowp(j))) = make_double2(w4 * (m3 * t0), w4 * (m3 * t1));
      }
    }
    __syncthreads();
    // ---- td2 = (W3^T td3)(1 - a2^2) + q2 ((-2 a2) ta2) -> A1 (in place)
    seed_acc(acc, tm, [](int) { return 0.0; });
    prefetch_act(actj, 1, tm, pa);
    prefetch_act(actj, 4, tm, pq);
    gemm_tile(A2, Wt.W3, WD, wbuf, tm, acc);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int c = tm.col(f);
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const double a2 = pa[j >> 1][f], q2 = pq[j >> 1][f];
        const double e2 = fma(-a2, a2, 1.0), m2 = -2.0 * a2;
        double2* p = reinterpret_cast<double2*>(A1 + aidx(c, tm.rowp(j)));
        const double2 ta2 = *p;
        *p = make_double2(fma(q2, m2 * ta2.x, acc[j][f][0] * e2), fma(q2, m2 * ta2.y, acc[j][f][1] * e2));
      }
    }
    __syncthreads();
    // ---- td1 = (W2^T td2)(1 - a1^2) + q1 ((-2 a1) ta1) -> A0 (in place)
    seed_acc(acc, tm, [](int) { return 0.0; });
    prefetch_act(actj, 0, tm, pa);
    prefetch_act(actj, 3, tm, pq);
    gemm_tile(A1, Wt.W2, WD, wbuf, tm, acc);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int c = tm.col(f);
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const double a1 = pa[j >> 1][f], q1 = pq[j >> 1][f];
        const double e1 = fma(-a1, a1, 1.0), m1 = -2.0 * a1;
        double2* p = reinterpret_cast<double2*>(A0 + aidx(c, tm.rowp(j)));
        const double2 ta1 = *p;
        *p = make_double2(fma(q1, m1 * ta1.x, acc[j][f][0] * e1), fma(q1, m1 * ta1.y, acc[j][f][1] * e1));
      }
    }
    __syncthreads();
    // ---- Hvv[jj][k] = sum_i W1[i][jj] td1[i][row(slot, k)];  J[r][c] = (r < N ? sign : 1) Hvv[r][vc(c)]
    {
      const OutTile ot;
      double g[OPT][2];
      output_product(A0, Wt.W1, tm, ot, g);
#pragma unroll
      for (int e2 = 0; e2 < 2; e2++) {
        const int r = ot.row0(tm) + e2;
        const int slot = r >> 4, k = r & 15;
        const int m = s_traj[slot];
        if (m >= 0) {
          const int cc = (k + NH) & (NV - 1);   // column of y whose v-index is k
#pragma unroll
          for (int t = 0; t < OPT; t++) {
            const int jj = ot.feature(t, tm);
            const double v = (jj < NH) ? io.sign * g[t][e2] : g[t][e2];
            io.J[(size_t)(jj * NV + cc) * io.J_se + (size_t)m * io.J_sr] = v;
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace mlp
}  // namespace jb
