"""TEST INFRASTRUCTURE (golden-vector generation only).  Stand-in for the un-vendored janus-linear Python module
`LUBatch` that /root/reference/src/python/seulext.py:3 imports: LUBatch(a[M,D,D]), .solvev(b[M,D]) and
.solvef(mask[M], b) as its call sites use them (seulext.py:597-599, 642-653), on torch.linalg (LAPACK getrf/getrs:
row partial pivoting, the same pivot rule as csrc/lu.cuh)."""
import torch


class LUBatch:
    def __init__(self, a):
        self.LU, self.piv = torch.linalg.lu_factor(a)

    def solvev(self, b):
        return torch.linalg.lu_solve(self.LU, self.piv, b.unsqueeze(-1)).squeeze(-1)

    def solvef(self, mask, b):
        LU, piv = self.LU[mask], self.piv[mask]
        rhs = b if b.shape[0] == LU.shape[0] else b[mask]
        return torch.linalg.lu_solve(LU, piv, rhs.unsqueeze(-1)).squeeze(-1)
