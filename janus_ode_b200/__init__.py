"""janus_ode_b200 -- B200-native batched Hamiltonian Radau integrator (one hot path of
optstat/janus-ode behind its solver API).  See DESIGN.md / INTEGRATION.md."""
from ._lib import JanusB200Error, LIB_PATH
from .hamiltonians import Hamiltonian, Harmonic, MlpHamiltonian, VanDerPol, VdpAug, VdpPmp
from .options import STAT_NAMES, LinearEvents, OptionsTe, StatsTe
from .user import compile_hamiltonian
from .solver import Dopri853Te, Dopri853TeD, SeulexState, SeulexT, SeulexTedState, SeulexTed, HamiltonianSystem, RadauState, RadauT, RadauTe, RadauTeD, det_pow, fp64_dmma_tflops, fp64_peak_tflops, lu_solve

__all__ = ["JanusB200Error", "LIB_PATH", "Hamiltonian", "Harmonic", "VanDerPol", "VdpAug", "VdpPmp", "MlpHamiltonian",
           "STAT_NAMES", "OptionsTe", "StatsTe", "HamiltonianSystem", "RadauState", "RadauT", "RadauTe", "det_pow",
           "fp64_peak_tflops", "fp64_dmma_tflops", "lu_solve", "compile_hamiltonian", "Dopri853Te", "Dopri853TeD", "RadauTeD", "SeulexState", "SeulexT", "SeulexTedState", "SeulexTed", "LinearEvents"]
