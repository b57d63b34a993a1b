// One translation unit per built-in Hamiltonian (-DJB_TU_HAM=<id>), so that the library builds in
// parallel: each TU instantiates the fused / epoch / stage-eval kernels for its Hamiltonian and
// exports one HamOps table.
#include "solver_impl.cuh"

namespace jb {
#if JB_TU_HAM == 0
extern const HamOps jb_ops_harmonic = make_ops<HamHarmonic>();
#elif JB_TU_HAM == 1
extern const HamOps jb_ops_vdp = make_ops<HamVdp>();
#elif JB_TU_HAM == 2
extern const HamOps jb_ops_vdp_pmp_smooth = make_ops<HamVdpPmp<false>>();
#elif JB_TU_HAM == 3
extern const HamOps jb_ops_vdp_pmp_bang = make_ops<HamVdpPmp<true>>();
#elif JB_TU_HAM == 4
extern const HamOps jb_ops_vdp_aug_bang = make_ops<HamVdpAugBang>();
#else
#error "JB_TU_HAM must be 0..4"
#endif
}  // namespace jb
