// Thread-local small dense LU with partial pivoting, real and complex (the janus-linear /
// torch::linalg_lu_factor + lu_solve replacement on the Radau path).
//
// Reference call sites: /root/reference/src/cpp/radaute_impl.hpp:1558,1611 (lu_factor of
// gamma/h I - J and (alpha + i beta)/h I - J, complex128) and :1754,1832,1927,1966 (lu_solve).
// One thread owns one n x n matrix; n is a template parameter so every loop unrolls and the
// matrix lives in registers (n <= 4) -- no shared-memory traffic, no shuffles.
//
// Pivot rule = LAPACK zgetf2 / izamax: first row with the largest |re| + |im| (strict >), rows
// physically exchanged, multipliers formed by multiplying with the reciprocal of the pivot.
// The row exchange is a predicated select over the candidate rows (register arrays cannot be
// indexed dynamically).  The reciprocals of U's diagonal are kept for the solves.
//
// ARITHMETIC CONTRACT (mirrored by oracle/c/radau_oracle.c "mirror" mode):
//   real     r = 1/a_kk ; l = a_ik*r ; a_ij = fma(-l, a_kj, a_ij)
//   complex  inv = 1/fma(d,d,c*c) ; rr = c*inv ; ri = -d*inv      (pivot c + i d)
//            lr = fma(-xi, ri, xr*rr) ; li = fma(xi, rr, xr*ri)
//            ar_ij = fma(li, ui, fma(-lr, ur, ar_ij)) ; ai_ij = fma(-li, ur, fma(-lr, ui, ai_ij))
//   solve    forward  s = fma(-l_ij, b_j, s)  (complex: sr = fma(li, bi_j, fma(-lr, br_j, sr)),
//                                                       si = fma(-li, br_j, fma(-lr, bi_j, si)))
//            backward same with u_ij, then b_i = s * rd_i
//            (complex: br = fma(-si, ri, sr*rr), bi = fma(si, rr, sr*ri))
#pragma once
#include "detmath.cuh"

namespace jb {

// JB_LU_BRANCH: skip a column's row exchange when this thread's pivot is the diagonal -- in a coherent warp (rows ordered by
// H(y0) / step size) the lanes mostly agree, so the whole select network of that column is skipped; a lane that does need
// the exchange runs the same selects as before (same arithmetic, same bits).  Measured: see profiles/r02_lu_branch_ab.txt
#if defined(JB_LU_BRANCH)
#define JB_LU_IF(c) if (c)
#else
#define JB_LU_IF(c)
#endif

template <int N>
struct RLU {
  double a[N * N];  // L\U row-major
  double rd[N];     // 1 / U_kk
  int piv[N];
};

template <int N>
struct ZLU {
  double re[N * N], im[N * N];
  double rdre[N], rdim[N];
  int piv[N];
};

template <int N>
JB_HD void rlu_factor(RLU<N>& L) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    int p = k;
    double best = fabs(L.a[k * N + k]);
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const double v = fabs(L.a[i * N + k]);
      if (v > best) { best = v; p = i; }
    }
    L.piv[k] = p;
    JB_LU_IF(p != k)
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const bool sw = (p == i);
#pragma unroll
      for (int j = 0; j < N; j++) {
        const double x = L.a[k * N + j], y = L.a[i * N + j];
        L.a[k * N + j] = sw ? y : x;
        L.a[i * N + j] = sw ? x : y;
      }
    }
    const double r = 1.0 / L.a[k * N + k];
    L.rd[k] = r;
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const double l = L.a[i * N + k] * r;
      L.a[i * N + k] = l;
#pragma unroll
      for (int j = k + 1; j < N; j++) L.a[i * N + j] = fma(-l, L.a[k * N + j], L.a[i * N + j]);
    }
  }
}

template <int N>
JB_HD void rlu_solve(const RLU<N>& L, double* b) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    JB_LU_IF(L.piv[k] != k)
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const bool sw = (L.piv[k] == i);
      const double x = b[k], y = b[i];
      b[k] = sw ? y : x;
      b[i] = sw ? x : y;
    }
  }
#pragma unroll
  for (int i = 1; i < N; i++) {
    double s = b[i];
#pragma unroll
    for (int j = 0; j < i; j++) s = fma(-L.a[i * N + j], b[j], s);
    b[i] = s;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double s = b[i];
#pragma unroll
    for (int j = i + 1; j < N; j++) s = fma(-L.a[i * N + j], b[j], s);
    b[i] = s * L.rd[i];
  }
}

template <int N>
JB_HD void zlu_factor(ZLU<N>& L) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    int p = k;
    double best = fabs(L.re[k * N + k]) + fabs(L.im[k * N + k]);
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const double v = fabs(L.re[i * N + k]) + fabs(L.im[i * N + k]);
      if (v > best) { best = v; p = i; }
    }
    L.piv[k] = p;
    JB_LU_IF(p != k)
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const bool sw = (p == i);
#pragma unroll
      for (int j = 0; j < N; j++) {
        double x = L.re[k * N + j], y = L.re[i * N + j];
        L.re[k * N + j] = sw ? y : x;
        L.re[i * N + j] = sw ? x : y;
        x = L.im[k * N + j];
        y = L.im[i * N + j];
        L.im[k * N + j] = sw ? y : x;
        L.im[i * N + j] = sw ? x : y;
      }
    }
    const double c = L.re[k * N + k], d = L.im[k * N + k];
    const double inv = 1.0 / fma(d, d, c * c);
    const double rr = c * inv, ri = -d * inv;
    L.rdre[k] = rr;
    L.rdim[k] = ri;
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const double xr = L.re[i * N + k], xi = L.im[i * N + k];
      const double lr = fma(-xi, ri, xr * rr), li = fma(xi, rr, xr * ri);
      L.re[i * N + k] = lr;
      L.im[i * N + k] = li;
#pragma unroll
      for (int j = k + 1; j < N; j++) {
        const double ur = L.re[k * N + j], ui = L.im[k * N + j];
        L.re[i * N + j] = fma(li, ui, fma(-lr, ur, L.re[i * N + j]));
        L.im[i * N + j] = fma(-li, ur, fma(-lr, ui, L.im[i * N + j]));
      }
    }
  }
}

template <int N>
JB_HD void zlu_solve(const ZLU<N>& L, double* br, double* bi) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    JB_LU_IF(L.piv[k] != k)
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const bool sw = (L.piv[k] == i);
      double x = br[k], y = br[i];
      br[k] = sw ? y : x;
      br[i] = sw ? x : y;
      x = bi[k];
      y = bi[i];
      bi[k] = sw ? y : x;
      bi[i] = sw ? x : y;
    }
  }
#pragma unroll
  for (int i = 1; i < N; i++) {
    double sr = br[i], si = bi[i];
#pragma unroll
    for (int j = 0; j < i; j++) {
      const double lr = L.re[i * N + j], li = L.im[i * N + j];
      sr = fma(li, bi[j], fma(-lr, br[j], sr));
      si = fma(-li, br[j], fma(-lr, bi[j], si));
    }
    br[i] = sr;
    bi[i] = si;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double sr = br[i], si = bi[i];
#pragma unroll
    for (int j = i + 1; j < N; j++) {
      const double ur = L.re[i * N + j], ui = L.im[i * N + j];
      sr = fma(ui, bi[j], fma(-ur, br[j], sr));
      si = fma(-ui, br[j], fma(-ur, bi[j], si));
    }
    const double rr = L.rdre[i], ri = L.rdim[i];
    br[i] = fma(-si, ri, sr * rr);
    bi[i] = fma(si, rr, sr * ri);
  }
}


// ------------------------------------------------------------------ complex factors in shared memory
// Optional placement (JB_LC_SMEM, device code): the complex L\U of a trajectory is written once
// per factorisation and read once per Newton iteration -- 40 doubles per block for n = 4, which
// the 5- and 7-stage steppers otherwise spill to local memory (more than the L1 holds at 256
// threads per SM).  Word e of block b of this thread lives at
//   jb_lc_smem[(b * WORDS + e) * blockDim.x + threadIdx.x]     (conflict-free),
// the pivots stay in registers.  Same arithmetic as zlu_solve, bit for bit.
template <int N>
struct ZLUS {
  static constexpr int WORDS = 2 * N * N + 2 * N;
  int piv[N];
  int base;   // b * WORDS
#if !(defined(__CUDA_ARCH__) && defined(JB_LC_SMEM))
  ZLU<N> host;   // host pass / placement disabled: a plain copy
#endif
};

#if defined(__CUDA_ARCH__) && defined(JB_LC_SMEM)
__device__ __forceinline__ double& lc_smem(int e) {
  extern __shared__ double jb_lc_smem[];
  return jb_lc_smem[e * blockDim.x + threadIdx.x];
}
#endif

template <int N>
JB_HD void zlus_store(ZLUS<N>& S, const ZLU<N>& L, int block) {
  S.base = block * ZLUS<N>::WORDS;
#pragma unroll
  for (int k = 0; k < N; k++) S.piv[k] = L.piv[k];
#if defined(__CUDA_ARCH__) && defined(JB_LC_SMEM)
#pragma unroll
  for (int k = 0; k < N * N; k++) {
    lc_smem(S.base + k) = L.re[k];
    lc_smem(S.base + N * N + k) = L.im[k];
  }
#pragma unroll
  for (int k = 0; k < N; k++) {
    lc_smem(S.base + 2 * N * N + k) = L.rdre[k];
    lc_smem(S.base + 2 * N * N + N + k) = L.rdim[k];
  }
#else
  S.host = L;
#endif
}

template <int N>
JB_HD void zlu_solve(const ZLUS<N>& S, double* br, double* bi) {
#if defined(__CUDA_ARCH__) && defined(JB_LC_SMEM)
  const int b0 = S.base;
#pragma unroll
  for (int k = 0; k < N; k++) {
    JB_LU_IF(S.piv[k] != k)
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const bool sw = (S.piv[k] == i);
      double x = br[k], y = br[i];
      br[k] = sw ? y : x;
      br[i] = sw ? x : y;
      x = bi[k];
      y = bi[i];
      bi[k] = sw ? y : x;
      bi[i] = sw ? x : y;
    }
  }
#pragma unroll
  for (int i = 1; i < N; i++) {
    double sr = br[i], si = bi[i];
#pragma unroll
    for (int j = 0; j < i; j++) {
      const double lr = lc_smem(b0 + i * N + j), li = lc_smem(b0 + N * N + i * N + j);
      sr = fma(li, bi[j], fma(-lr, br[j], sr));
      si = fma(-li, br[j], fma(-lr, bi[j], si));
    }
    br[i] = sr;
    bi[i] = si;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double sr = br[i], si = bi[i];
#pragma unroll
    for (int j = i + 1; j < N; j++) {
      const double ur = lc_smem(b0 + i * N + j), ui = lc_smem(b0 + N * N + i * N + j);
      sr = fma(ui, bi[j], fma(-ur, br[j], sr));
      si = fma(-ui, br[j], fma(-ur, bi[j], si));
    }
    const double rr = lc_smem(b0 + 2 * N * N + i), ri = lc_smem(b0 + 2 * N * N + N + i);
    br[i] = fma(-si, ri, sr * rr);
    bi[i] = fma(si, rr, sr * ri);
  }
#else
  zlu_solve<N>(S.host, br, bi);
#endif
}

// ------------------------------------------------------------------ compact complex factors, runtime block index
// Storage of the rolled high-order steppers (radau_core.cuh: attempt_step_hi): ALL (s-1)/2 complex blocks of a
// trajectory live in shared memory and are addressed by a RUN-TIME block number, so the factor / solve code exists
// once in the instruction stream instead of once per block (the unrolled 5- and 7-stage steppers were 186 / 255 kB
// of SASS and instruction-fetch bound: ncu `no_instruction` 4.3 / 5.4 of ~8 stall cycles per issue, round-2 profile).
// Only what the solves read is kept: the strictly lower / upper parts (N^2 - N complex numbers), the reciprocals of
// U's diagonal (N) and the pivots packed into one word: 2 N^2 + 1 words per block (33 for N = 4; three blocks of a
// 7-stage stepper at 2 x 128 threads per SM = 198 kB of the 227 kB).  Same arithmetic as zlu_solve, bit for bit.
template <int N>
struct ZLUC {
  static constexpr int OFF = N * N - N;
  static constexpr int WORDS = 2 * N * N + 1;
  JB_HD static constexpr int idx(int i, int j) { return i * (N - 1) + (j < i ? j : j - 1); }
};

#if defined(__CUDA_ARCH__) && defined(JB_LC_SMEM)
template <int N>
__device__ __forceinline__ void zluc_store(int block, const ZLU<N>& L) {
  const int b0 = block * ZLUC<N>::WORDS;
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++)
      if (i != j) {
        lc_smem(b0 + ZLUC<N>::idx(i, j)) = L.re[i * N + j];
        lc_smem(b0 + ZLUC<N>::OFF + ZLUC<N>::idx(i, j)) = L.im[i * N + j];
      }
  int pk = 0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    lc_smem(b0 + 2 * ZLUC<N>::OFF + k) = L.rdre[k];
    lc_smem(b0 + 2 * ZLUC<N>::OFF + N + k) = L.rdim[k];
    pk |= L.piv[k] << (4 * k);
  }
  lc_smem(b0 + 2 * N * N) = __longlong_as_double((long long)pk);
}

template <int N>
__device__ __forceinline__ void zluc_solve(int block, double* br, double* bi) {
  const int b0 = block * ZLUC<N>::WORDS;
  const int pk = (int)__double_as_longlong(lc_smem(b0 + 2 * N * N));
#pragma unroll
  for (int k = 0; k < N; k++) {
    const int pv = (pk >> (4 * k)) & 15;
#pragma unroll
    for (int i = k + 1; i < N; i++) {
      const bool sw = (pv == i);
      double x = br[k], y = br[i];
      br[k] = sw ? y : x;
      br[i] = sw ? x : y;
      x = bi[k];
      y = bi[i];
      bi[k] = sw ? y : x;
      bi[i] = sw ? x : y;
    }
  }
#pragma unroll
  for (int i = 1; i < N; i++) {
    double sr = br[i], si = bi[i];
#pragma unroll
    for (int j = 0; j < i; j++) {
      const double lr = lc_smem(b0 + ZLUC<N>::idx(i, j)), li = lc_smem(b0 + ZLUC<N>::OFF + ZLUC<N>::idx(i, j));
      sr = fma(li, bi[j], fma(-lr, br[j], sr));
      si = fma(-li, br[j], fma(-lr, bi[j], si));
    }
    br[i] = sr;
    bi[i] = si;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double sr = br[i], si = bi[i];
#pragma unroll
    for (int j = i + 1; j < N; j++) {
      const double ur = lc_smem(b0 + ZLUC<N>::idx(i, j)), ui = lc_smem(b0 + ZLUC<N>::OFF + ZLUC<N>::idx(i, j));
      sr = fma(ui, bi[j], fma(-ur, br[j], sr));
      si = fma(-ui, br[j], fma(-ur, bi[j], si));
    }
    const double rr = lc_smem(b0 + 2 * ZLUC<N>::OFF + i), ri = lc_smem(b0 + 2 * ZLUC<N>::OFF + N + i);
    br[i] = fma(-si, ri, sr * rr);
    bi[i] = fma(si, rr, sr * ri);
  }
}
#endif

// which steppers run rolled (all complex blocks in shared memory, z / w in local memory): 5 and 7 stages on the device.
// OPT-IN (-DJB_ROLL_HI).  Measured (round 2, config 2, 2^20 rows): the rolled 5/7-stage kernels are 39 % smaller and 7 %
// faster on the epochs they dominate when every epoch is synchronised (89.6 vs 91.0 ms per solve), but in the real
// pipeline -- the three per-order kernels of an epoch overlapping on three streams -- the solve is 5 % SLOWER (95.5 vs
// 90.2 ms, bench.py, twice): their larger shared-memory footprint (99 instead of 80 words per thread) no longer lets a
// 3-stage CTA share an SM with two 7-stage CTAs, which costs more overlap than the smaller code gains.
template <int N, int S>
struct HiRoll {
#if defined(JB_LC_SMEM) && defined(JB_ROLL_HI)
#ifndef JB_ROLL_HI_MIN
#define JB_ROLL_HI_MIN 5
#endif
  static constexpr bool on = (S >= JB_ROLL_HI_MIN) && (N <= 4);
#else
  static constexpr bool on = false;
#endif
};

// how many of the (s-1)/2 complex blocks of an s-stage stepper are parked in shared memory
template <int N, int S>
struct LcPlace {
#if defined(JB_LC_SMEM)
  static constexpr int NQ = (S - 1) / 2;
  static constexpr int NS = (N <= 4) ? (NQ < 2 ? NQ : 2) : 0;   // 40 words per block and thread for n = 4
#else
  static constexpr int NQ = (S - 1) / 2;
  static constexpr int NS = 0;
#endif
  static constexpr int NR = NQ - NS;
};

}  // namespace jb
