// Shapes and algorithmic FLOP counts of the MLP Hamiltonian path (config 4), shared by the tile
// kernels (mlp_kernels.cuh) and the phase functions (mlp_phase.cuh, also compiled for the host).
#pragma once

namespace jb {
namespace mlp {

constexpr int NV = 16;        // inputs of H: [x(8) ; p(8)]
constexpr int NH = 8;         // N = dim x = dim p
constexpr int WD = 128;       // hidden width

// algorithmic FLOPs (SURVEY.md 8d convention: mul, add, tanh = 1) of one gradient / one Jacobian
constexpr long long MACS = (long long)NV * WD + 2LL * WD * WD;          // 34,816 multiply-adds / sweep
constexpr long long FLOPS_F = 2 * (2 * MACS) + 3LL * WD * 4;           // forward + reverse + elementwise
constexpr long long FLOPS_J = (long long)NV * (2 * (2LL * 2 * WD * WD) + 2LL * WD * NV + 16LL * WD);

}  // namespace mlp
}  // namespace jb
