// 16 x 16 LU with partial pivoting by 16 lanes (half a warp) per matrix, one row per lane, the
// row held in registers; real and complex (the janus-linear / torch::linalg_lu_factor replacement
// for Ny = 16, /root/reference/src/cpp/radaute_impl.hpp:1558,1611).
//
// Same operations in the same order as the scalar rlu_factor_g / zlu_factor_g (mlp_phase.cuh) --
// pivot = first row with the largest |re| + |im| (LAPACK izamax rule), rows physically exchanged,
// multipliers by the reciprocal of the pivot, a_ij = fma chain -- so the factors are bit-identical;
// only the ownership of the work changes: the pivot search is a 4-step shuffle reduction, the row
// exchange and the pivot-row broadcast are shuffles, every lane below the pivot eliminates its row.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

namespace jb {
namespace mlp {

// index of the first maximum over the 16 lanes of this half-warp (candidates: lanes >= k)
__device__ __forceinline__ int half_argmax(unsigned mask, double v, int l, int k) {
  // scalar rule: start with row k, replace by a later row only if strictly greater; NaN never wins
  // over a number, and a NaN in row k keeps row k
  double key = (l < k) ? -1.0 : v;
  if (key != key) key = (l == k) ? __longlong_as_double(0x7ff0000000000000LL) : -1.0;
  int idx = l;
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const double ok = __shfl_xor_sync(mask, key, off, 16);
    const int oi = __shfl_xor_sync(mask, idx, off, 16);
    if (ok > key || (ok == key && oi < idx)) {
      key = ok;
      idx = oi;
    }
  }
  return idx;
}

// compile-time loop 0 .. N-1: the column index of a factorisation step must be a constant for the row arrays to stay in
// registers -- with `#pragma unroll` the compiler unrolled the first columns and left a rolled remainder, whose run-time
// index put the rows in local memory
template <int K, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (K < N) {
    f(std::integral_constant<int, K>{});
    static_for<K + 1, N>(f);
  }
}

// real: lane l holds row l in a[16]; on return a = row l of L\U, rd = 1/U_ll, piv = pivot row of step l
__device__ __forceinline__ void rlu16_factor(unsigned mask, int l, double (&a)[16], double& rd, int& piv) {
  static_for<0, 16>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    const int p = half_argmax(mask, fabs(a[k]), l, k);
    if (l == k) piv = p;
    if (__any_sync(mask, p != k)) {
      const int src = (l == k) ? p : (l == p) ? k : l;
#pragma unroll
      for (int j = 0; j < 16; j++) a[j] = __shfl_sync(mask, a[j], src, 16);
    }
    const double r = 1.0 / __shfl_sync(mask, a[k], k, 16);
    if (l == k) rd = r;
    const bool below = l > k;
    const double lm = a[k] * r;
    if (below) a[k] = lm;
#pragma unroll
    for (int j = k + 1; j < 16; j++) {
      const double u = __shfl_sync(mask, a[j], k, 16);
      if (below) a[j] = fma(-lm, u, a[j]);
    }
  });
}

__device__ __forceinline__ void zlu16_factor(unsigned mask, int l, double (&re)[16], double (&im)[16], double& rdre,
                                             double& rdim, int& piv) {
  static_for<0, 16>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    const int p = half_argmax(mask, fabs(re[k]) + fabs(im[k]), l, k);
    if (l == k) piv = p;
    if (__any_sync(mask, p != k)) {
      const int src = (l == k) ? p : (l == p) ? k : l;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        re[j] = __shfl_sync(mask, re[j], src, 16);
        im[j] = __shfl_sync(mask, im[j], src, 16);
      }
    }
    const double c = __shfl_sync(mask, re[k], k, 16), d = __shfl_sync(mask, im[k], k, 16);
    const double inv = 1.0 / fma(d, d, c * c);
    const double rr = c * inv, ri = -d * inv;
    if (l == k) {
      rdre = rr;
      rdim = ri;
    }
    const bool below = l > k;
    const double xr = re[k], xi = im[k];
    const double lr = fma(-xi, ri, xr * rr), li = fma(xi, rr, xr * ri);
    if (below) {
      re[k] = lr;
      im[k] = li;
    }
#pragma unroll
    for (int j = k + 1; j < 16; j++) {
      const double ur = __shfl_sync(mask, re[j], k, 16), ui = __shfl_sync(mask, im[j], k, 16);
      if (below) {
        re[j] = fma(li, ui, fma(-lr, ur, re[j]));
        im[j] = fma(-li, ur, fma(-lr, ui, im[j]));
      }
    }
  });
}


// ---- solves with the factors above: lane l holds row l of L\U, 1/U_ll, the pivot of step l and
// element l of the right-hand side.  Column-oriented substitution, one broadcast + one fma per lane
// and column: forward  b_i -= l_ij b_j  (j = 0..14, rows i > j: row i accumulates in increasing j,
// as the scalar code), backward  x_j = b_j / u_jj, b_i -= u_ij x_j  (j = 15..0, rows i < j) --
// the same sequences as rlu_solve_g / zlu_solve_g (mlp_phase.cuh).
__device__ __forceinline__ void rlu16_solve(unsigned mask, int l, const double (&a)[16], double rd, int piv, double& b) {
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int p = __shfl_sync(mask, piv, k, 16);
    const int src = (l == k) ? p : (l == p) ? k : l;
    b = __shfl_sync(mask, b, src, 16);
  }
#pragma unroll
  for (int j = 0; j < 15; j++) {
    const double bj = __shfl_sync(mask, b, j, 16);
    if (l > j) b = fma(-a[j], bj, b);
  }
#pragma unroll
  for (int j = 15; j >= 0; j--) {
    const double xj = __shfl_sync(mask, b * rd, j, 16);
    if (l == j) b = xj;
    if (l < j) b = fma(-a[j], xj, b);
  }
}

__device__ __forceinline__ void zlu16_solve(unsigned mask, int l, const double (&re)[16], const double (&im)[16],
                                            double rdre, double rdim, int piv, double& br, double& bi) {
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int p = __shfl_sync(mask, piv, k, 16);
    const int src = (l == k) ? p : (l == p) ? k : l;
    br = __shfl_sync(mask, br, src, 16);
    bi = __shfl_sync(mask, bi, src, 16);
  }
#pragma unroll
  for (int j = 0; j < 15; j++) {
    const double xr = __shfl_sync(mask, br, j, 16), xi = __shfl_sync(mask, bi, j, 16);
    if (l > j) {
      const double nr = fma(im[j], xi, fma(-re[j], xr, br));
      bi = fma(-im[j], xr, fma(-re[j], xi, bi));
      br = nr;
    }
  }
#pragma unroll
  for (int j = 15; j >= 0; j--) {
    const double qr = fma(-bi, rdim, br * rdre), qi = fma(bi, rdre, br * rdim);
    const double xr = __shfl_sync(mask, qr, j, 16), xi = __shfl_sync(mask, qi, j, 16);
    if (l == j) {
      br = xr;
      bi = xi;
    }
    if (l < j) {
      const double nr = fma(im[j], xi, fma(-re[j], xr, br));
      bi = fma(-im[j], xr, fma(-re[j], xi, bi));
      br = nr;
    }
  }
}

}  // namespace mlp
}  // namespace jb
