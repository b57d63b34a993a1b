// TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
// Stand-in for the un-vendored janus-tensor-dual header <janus/tensordual.hpp>, which the reference's
// src/cpp/radaute.hpp:16 and src/cpp/janus_ode_common.hpp:4 include.  RadauTe itself (radaute_impl.hpp)
// only needs the type name (radaute.hpp:40, an unused typedef), `bpow` (:439,569,628,... = elementwise
// power) and `print_tensor` (:1838-1849, debug dumps in dead branches).  The templates evalDynsDual /
// evalJacDual (janus_ode_common.hpp:2075-2428) additionally need value+dual containers with `r`, `d`,
// `index` and `cat`; they are written here from how those call sites use them.
#pragma once
#include <torch/torch.h>
#include <iostream>
#include <vector>

namespace janus {

struct TensorDual {
  torch::Tensor r, d;   // r [M,N], d [M,N,D]
  TensorDual() {}
  TensorDual(const torch::Tensor& r_, const torch::Tensor& d_) : r(r_), d(d_) {}
  TensorDual index(const std::vector<torch::indexing::TensorIndex>& ix) const {
    return TensorDual(r.index(ix), d.index(ix));
  }
  static TensorDual cat(const std::vector<TensorDual>& parts) {
    std::vector<torch::Tensor> rs, ds;
    for (auto& p : parts) { rs.push_back(p.r); ds.push_back(p.d); }
    return TensorDual(torch::cat(rs, 1), torch::cat(ds, 1));
  }
};

struct TensorMatDual {
  torch::Tensor r, d;   // r [M,N,L], d [M,N,L,D]
  TensorMatDual() {}
  TensorMatDual(const torch::Tensor& r_, const torch::Tensor& d_) : r(r_), d(d_) {}
  static TensorMatDual cat(const TensorMatDual& a, const TensorMatDual& b, int dim) {
    return TensorMatDual(torch::cat({a.r, b.r}, dim), torch::cat({a.d, b.d}, dim));
  }
};

inline torch::Tensor bpow(const torch::Tensor& a, const torch::Tensor& b) { return torch::pow(a, b); }
inline torch::Tensor bpow(const torch::Tensor& a, const torch::Scalar& b) { return torch::pow(a, b); }
inline void print_tensor(const torch::Tensor& a) { std::cerr << a << std::endl; }

}  // namespace janus
