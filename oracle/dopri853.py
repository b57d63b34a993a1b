"""ORACLE (test infrastructure only): numpy restatement of the reference's explicit stepper
/root/reference/src/python/dopri853te.py -- Dopri853Te.step(htry): dy :207-287, error :191-205, control :318-355,
step :290-315 -- one accepted step of every trajectory per call, rejected attempts retried.

Pinned against the LIVE reference: tests/golden/make_dopri_golden.py imports the reference class in the build
container, runs it on seeded batches and stores every step; tests/test_dopri853.py checks this restatement
against those vectors (agreement to roundoff: the reference sums with torch's vectorised reduction and calls
torch.pow, this file sums left to right and uses three square roots for err^(-1/8)).

Reference quirks reproduced (see csrc/dopri853.cuh): float32-rounded coefficients / safe / minscale, error norm
divided by D * M with M the batch size, hnext = h * min(scale, 1), beta = 0, and the y / yout aliasing:
step() ends with `self.y = self.yout` (:311), so from the SECOND call on both names are one tensor -- the
attempt's result overwrites y before error() reads it (sk = atol + rtol |yout| only) and a rejected attempt leaves
its rejected yout behind as the state the retry starts from (with the old dydx)."""
import numpy as np

from .dopri853_tables import ALPHA, MAXSCALE, MINSCALE, REF, SAFE

EPS = np.finfo(np.float64).eps


class Dopri853Oracle:
    def __init__(self, y, dydx, x, atol, rtol, derivs, norm_M=None, tables=REF):
        self.y = np.array(y, dtype=np.float64)
        self.dydx = np.array(dydx, dtype=np.float64)
        self.x = np.array(x, dtype=np.float64)
        self.M, self.D = self.y.shape
        self.atol, self.rtol, self.derivs = float(atol), float(rtol), derivs
        self.norm_M = int(norm_M or self.M)
        self.errold = np.full(self.M, 1.0e-4)
        self.hnext = np.zeros(self.M)
        self.hdid = np.zeros(self.M)
        self.nrej = np.zeros(self.M, dtype=np.int32)
        self.tab = tables
        self.aliased = False          # y is yout (reference :311) after the first step()
        assert ALPHA == 0.125

    def _attempt(self, h, x, y, dydx):
        """dy() + error(): rows = the trajectories still rejected.  Returns yout, err."""
        aliased = self.aliased
        A, B, BHH = self.tab["A"], self.tab["B"], self.tab["BHH"]
        K = {1: dydx}
        he = h[:, None]
        for i in range(2, 13):                                   # stages 2..12 (time arguments: autonomous dynamics)
            acc = None
            for j in sorted(A[i]):                               # left to right, as written in the reference
                term = A[i][j] * K[j]
                acc = term if acc is None else acc + term
            K[i] = self.derivs(None, y + he * acc)
        kk = None
        for j in sorted(B):
            term = B[j] * K[j]
            kk = term if kk is None else kk + term
        yout = y + he * kk
        yerr = kk - BHH[1] * dydx - BHH[2] * K[9] - BHH[3] * K[12]
        yerr2 = K[5] - BHH[1] * dydx - BHH[2] * K[9] - BHH[3] * K[12]
        sk = self.atol + self.rtol * np.maximum(np.abs(yout if aliased else y), np.abs(yout))
        s1 = np.zeros(len(h))
        s2 = np.zeros(len(h))
        for i in range(self.D):                                  # left-to-right sums (what the GPU does)
            q1, q2 = yerr[:, i] / sk[:, i], yerr2[:, i] / sk[:, i]
            s1 = s1 + q1 * q1
            s2 = s2 + q2 * q2
        deno = s2 + 0.01 * s1
        deno = np.where(deno <= 0.0, 1.0, deno)
        err = np.abs(h) * s2 * np.sqrt(1.0 / (float(self.D * self.norm_M) * deno))
        return yout, err

    def step(self, htry):
        h = np.array(np.broadcast_to(np.asarray(htry, dtype=np.float64), (self.M,)))
        reject = h > EPS
        started = reject.copy()
        yout = self.y.copy()
        self.nrej[:] = 0
        while reject.any():
            m = np.nonzero(reject)[0]
            yo, err = self._attempt(h[m], self.x[m], self.y[m], self.dydx[m])
            yout[m] = yo
            root = 1.0 / np.sqrt(np.sqrt(np.sqrt(err)))          # err^-alpha, alpha = 1/8
            ok = err <= 1.0
            # accepted rows
            scale = np.clip(SAFE * root, MINSCALE, MAXSCALE)
            scale = np.where(self.errold[m] == 0.0, MAXSCALE, scale)
            a = m[ok]
            self.hnext[a] = h[a] * np.minimum(scale[ok], 1.0)
            self.errold[a] = np.maximum(err[ok], 1.0e-4)
            reject[a] = False
            # rejected rows
            r = m[~ok]
            h[r] = h[r] * np.maximum(SAFE * root[~ok], MINSCALE)
            self.nrej[r] += 1
            if self.aliased:
                self.y[r] = yout[r]                              # the rejected result has overwritten the state
        s = np.nonzero(started)[0]
        self.dydx[s] = self.derivs(None, yout[s])
        self.y[s] = yout[s]
        self.x[s] = self.x[s] + h[s]
        self.hdid = h
        self.aliased = True
        return self
