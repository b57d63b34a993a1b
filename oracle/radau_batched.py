"""Batched (masked-tensor) CPU restatement of the reference's RadauTe path -- TEST / BASELINE INFRASTRUCTURE.

Same op structure as the reference (/root/reference/src/cpp/radaute_impl.hpp:537-1345): every pass
attempts one step for all active samples with boolean masks, per-stage-count groups are gathered with
index ops, DecomRC / Solvrad / Estrad call torch.linalg.lu_factor / lu_solve on complex128 batches
(:1558,1611,1754,1832,1927) and the dynamics / Jacobian come from torch.autograd on the Hamiltonian
(oracle/ham.py == src/cpp/janus_ode_common.hpp:2161-2299).  This is the stand-in for "the reference's
own CPU implementation" that BASELINE.md section 3 prescribes as the timed CPU baseline (the reference
itself cannot be built here), and a third, independently written cross-check of the algorithm.

Semantics: per-trajectory (RadauTeD `any()` / MATLAB), same deviations Q3,Q4,Q6-Q9 as
oracle/radau_scalar.py; counters with MATLAB rules.  Only tests/ and bench.py import this module.
"""
from __future__ import annotations

import torch

from .radau_tables import TABLES

EPS = 2.220446049250313e-16
ST = dict(Fcn=0, Jac=1, Dec=2, Sol=3, Step=4, Acc=5, Rej=6, NRej=7)


def _tab(s):
    t = TABLES[s]
    f = lambda k: torch.tensor(t[k], dtype=torch.float64)
    return f("T"), f("TI"), f("C"), f("ValP"), f("Dd")


def _nit(s, mx):
    return {1: mx - 3, 3: mx, 5: mx + 5, 7: mx + 10}[s]


class RadauBatched:
    """RadauBatched(ode, jac, tspan[M,2], y0[M,Ny], options).solve() -> res[M]; fields y, t, h, stats[M,8], nstg.

    ode(t[m], y[m,Ny]) -> [m,Ny]; jac(t[m], y[m,Ny]) -> [m,Ny,Ny]; both are called with the masked subset,
    like the reference's OdeFcn/JacFn (src/cpp/radaute_impl.hpp:584,868-870)."""

    def __init__(self, ode, jac, tspan, y0, op):
        self.ode, self.jac, self.op = ode, jac, op
        self.y = y0.clone().double()
        self.M, self.Ny = self.y.shape
        M, Ny = self.M, self.Ny
        self.t = tspan[:, 0].clone().double()
        self.tf = tspan[:, -1].clone().double()
        self.sgn = torch.sign(self.tf - self.t)
        if bool((self.sgn == 0).any()):
            raise ValueError("Time vector must be strictly monotonic")
        self.RelTol = torch.as_tensor(op.RelTol, dtype=torch.float64).reshape(-1).expand(Ny).clone()
        self.AbsTol = torch.as_tensor(op.AbsTol, dtype=torch.float64).reshape(-1).expand(Ny).clone()
        self.hmaxn = (self.tf - self.t).abs()
        h = torch.full((M,), float(op.InitialStep), dtype=torch.float64)
        h[h.abs() <= 10 * EPS] = 1e-6
        self.h = self.sgn * torch.minimum(h.abs(), self.hmaxn)
        self.h_old = self.h.clone()
        self.hopt = self.h.clone()
        last = (self.t + self.h * 1.0001 - self.tf) * self.sgn >= 0
        self.h = torch.where(last, self.tf - self.t, self.h)
        self.Last = last.clone()
        z = lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt)
        self.FacConv = torch.ones(M, dtype=torch.float64)
        self.Theta, self.Thetat, self.hacc, self.erracc = z(M), z(M), z(M), z(M)
        self.NbrStg = torch.full((M,), int(op.NbrStg), dtype=torch.int64)
        self.Newt, self.ChangeNbr = z(M, dt=torch.int64), z(M, dt=torch.int64)
        b = lambda v: torch.full((M,), v, dtype=torch.bool)
        self.First, self.Reject, self.Keep, self.ChangeFlag = b(True), b(False), b(False), b(False)
        self.UnExpStepRej, self.UnExpNewtRej = b(False), b(False)
        self.NeedNewJac, self.NeedNewLU = b(True), b(True)
        self.Jac = z(M, Ny, Ny)
        self.cont = z(M, 7, Ny)                       # [sample, stage, component]
        self.LU = torch.zeros(M, 4, Ny, Ny, dtype=torch.complex128)
        self.piv = torch.zeros(M, 4, Ny, dtype=torch.int32)
        self.stats = z(M, 8, dt=torch.int64)
        self.res = z(M, dt=torch.int64)
        self.active = b(True)
        self.f0 = self.ode(self.t, self.y)
        self.stats[:, ST["Fcn"]] += 1
        nanrow = torch.isnan(self.f0).any(1)
        self.res[nanrow] = 2
        self.active &= ~nanrow
        self.count = 0

    # ------------------------------------------------------------------ linear algebra
    def _decomrc(self, idx, s, ValP):
        Ny = self.Ny
        valp = ValP[None, :] * (1.0 / self.h[idx])[:, None]                # einsum("s,m->ms", ValP, 1/h)
        eye = torch.eye(Ny, dtype=torch.float64)
        J = self.Jac[idx]
        B0 = -J + eye[None] * valp[:, 0, None, None]
        lu, pv = torch.linalg.lu_factor(torch.complex(B0, torch.zeros_like(B0)))
        self.LU[idx, 0], self.piv[idx, 0] = lu, pv
        for q in range(1, (s - 1) // 2 + 1):
            Br = eye[None] * valp[:, 2 * q - 1, None, None] - J
            Bi = eye[None] * valp[:, 2 * q, None, None]
            lu, pv = torch.linalg.lu_factor(torch.complex(Br, Bi.expand_as(Br).contiguous()))
            self.LU[idx, q], self.piv[idx, q] = lu, pv

    def _lusolve(self, idx, blk, rhs):
        return torch.linalg.lu_solve(self.LU[idx, blk], self.piv[idx, blk], rhs.unsqueeze(-1)).squeeze(-1)

    def _solvrad(self, idx, s, ValP, z, w):
        valp = ValP[None, :] * (1.0 / self.h[idx])[:, None]
        z = z.clone()
        z0 = z[:, 0] - valp[:, 0, None] * w[:, 0]
        z[:, 0] = self._lusolve(idx, 0, torch.complex(z0, torch.zeros_like(z0))).real
        for q in range(1, (s - 1) // 2 + 1):
            a, b = 2 * q - 1, 2 * q
            z2 = z[:, a] - valp[:, a, None] * w[:, a] + valp[:, b, None] * w[:, b]
            z3 = z[:, b] - valp[:, b, None] * w[:, a] - valp[:, a, None] * w[:, b]
            sol = self._lusolve(idx, q, torch.complex(z2, z3))
            z[:, a], z[:, b] = sol.real, sol.imag
        return z

    # ------------------------------------------------------------------ one pass for one stage group
    def _group_pass(self, idx, s):
        op, Ny = self.op, self.Ny
        T, TI, C, ValP, Dd = _tab(s)
        Nit = _nit(s, op.MaxNbrNewton)
        Variab = op.MaxNbrStg != op.MinNbrStg
        # DecomRC for the samples that need it
        m = self.NeedNewLU[idx]
        if bool(m.any()):
            sub = idx[m]
            self._decomrc(sub, s, ValP)
            self.stats[sub, ST["Dec"]] += 1
            self.NeedNewLU[sub] = False
        if Variab:
            k = idx[self.Keep[idx]]
            self.Keep[k] = False
            self.ChangeNbr[k] += 1
        # step size too small
        small = 0.1 * self.h[idx].abs() <= self.t[idx].abs() * EPS
        if bool(small.any()):
            bad = idx[small]
            self.res[bad] = 1
            self.active[bad] = False
            idx = idx[~small]
            if idx.numel() == 0:
                return
        ExpmNs = (s + 1.0) / (2.0 * s)
        RelTol1 = 0.1 * self.RelTol ** ExpmNs
        AbsTol1 = RelTol1 * (self.AbsTol / self.RelTol)
        Scal = AbsTol1[None] + RelTol1[None] * self.y[idx].abs()
        n = idx.numel()
        z = torch.zeros(n, s, Ny, dtype=torch.float64)
        w = torch.zeros(n, s, Ny, dtype=torch.float64)
        fresh = self.First[idx] | bool(op.Start_Newt) | self.ChangeFlag[idx]
        if bool(fresh.any()):
            self.cont[idx[fresh]] = 0.0
        old = ~fresh
        if bool(old.any()):
            io = idx[old]
            hq = self.h[io] / self.h_old[io]
            cq = C[None, :] * hq[:, None]
            cont = self.cont[io]
            zz = torch.empty(io.numel(), s, Ny, dtype=torch.float64)
            for q in range(s):
                acc = (cq[:, q] - C[0] + 1.0)[:, None] * cont[:, s - 1]
                for q1 in range(2, s + 1):
                    acc = (acc + cont[:, s - q1]) * (cq[:, q] - C[q1 - 1] + 1.0)[:, None]
                zz[:, q] = acc
            ww = torch.einsum("jq,mqn->mjn", TI, zz)
            z[old], w[old] = zz, ww
        mrt = float(RelTol1.min())
        FNewt = max(10 * EPS / mrt, min(0.03, mrt ** (1.0 / ExpmNs - 1.0)))
        if s == 1:
            FNewt = max(10 * EPS / mrt, 0.03)
        self.FacConv[idx] = torch.clamp(self.FacConv[idx], min=EPS) ** 0.8
        self.Theta[idx] = abs(op.JacRecompute)
        self.Newt[idx] = 0
        # ------- simplified Newton, masked (radaute_impl.hpp:839-985)
        it = torch.ones(n, dtype=torch.bool)         # still iterating
        failed = torch.zeros(n, dtype=torch.bool)    # Newton gave up -> skip the controller this pass
        OldNrm = torch.zeros(n, dtype=torch.float64)
        thqold = torch.zeros(n, dtype=torch.float64)
        NewNrm = torch.zeros(n, dtype=torch.float64)
        sq = (s * Ny) ** 0.5
        while bool(it.any()):
            loc = it.nonzero().squeeze(1)
            g = idx[loc]
            self.Reject[g] = False
            self.Newt[g] += 1
            over = self.Newt[g] > Nit
            if bool(over.any()):
                o_l, o_g = loc[over], g[over]
                self.UnExpStepRej[o_g] = True
                self.stats[o_g, ST["Rej"]] += 1
                self.stats[o_g, ST["NRej"]] += 1
                self.h[o_g] = self.h[o_g] * 0.5
                self.Reject[o_g] = True
                self.Last[o_g] = False
                self.NeedNewLU[o_g] = True
                failed[o_l] = True
                it[o_l] = False
                loc, g = loc[~over], g[~over]
                if loc.numel() == 0:
                    break
            fq = torch.empty(loc.numel(), s, Ny, dtype=torch.float64)
            for q in range(s):
                fq[:, q] = self.ode(self.t[g] + C[q] * self.h[g], self.y[g] + z[loc, q])
            nanf = torch.isnan(fq).flatten(1).any(1)
            if bool(nanf.any()):
                self.res[g[nanf]] = 2
                self.active[g[nanf]] = False
                failed[loc[nanf]] = True
                it[loc[nanf]] = False
                keep = ~nanf
                loc, g, fq = loc[keep], g[keep], fq[keep]
                if loc.numel() == 0:
                    break
            self.stats[g, ST["Fcn"]] += s
            zz = torch.einsum("jq,mqn->mjn", TI, fq)
            zz = self._solvrad(g, s, ValP, zz, w[loc])
            self.stats[g, ST["Sol"]] += 1
            nn = (zz / Scal[loc][:, None, :]).norm(dim=2).sum(1) / sq
            NewNrm[loc] = nn
            badn = ~torch.isfinite(nn)
            newt = self.Newt[g]
            chk = (newt > 1) & (newt < Nit) & ~badn
            stop = badn.clone()
            if bool(badn.any()):
                self.res[g[badn]] = 2
                self.active[g[badn]] = False
                failed[loc[badn]] = True
            if bool(chk.any()):
                cl, cg = loc[chk], g[chk]
                thq = nn[chk] / OldNrm[cl]
                th = torch.where(newt[chk] == 2, thq, torch.sqrt(thq * thqold[cl]))
                self.Theta[cg] = th
                thqold[cl] = thq
                conv = th < 0.99
                fc = th / (1.0 - th)
                self.FacConv[cg[conv]] = fc[conv]
                dyth = fc * nn[chk] * th ** (Nit - 1 - newt[chk]).double() / FNewt
                slow = conv & (dyth >= 1)
                if bool(slow.any()):
                    sg, sl = cg[slow], cl[slow]
                    qn = torch.clamp(dyth[slow], 1.0e-4, 20.0)
                    hh = 0.8 * qn ** (-1.0 / (4.0 + Nit - 1 - newt[chk][slow]).double())
                    self.h[sg] = hh * self.h[sg]
                    self.Reject[sg] = True
                    self.Last[sg] = False
                    self.UnExpNewtRej[sg] = hh <= 0.5
                    self.stats[sg, ST["NRej"]] += 1
                    self.stats[sg, ST["Rej"]] += 1
                    self.NeedNewLU[sg] = True
                    failed[sl] = True
                div = ~conv
                if bool(div.any()):
                    dg, dl = cg[div], cl[div]
                    self.h[dg] = self.h[dg] * 0.5
                    self.Reject[dg] = True
                    self.Last[dg] = False
                    self.UnExpStepRej[dg] = True
                    self.stats[dg, ST["Rej"]] += 1
                    self.NeedNewLU[dg] = True
                    failed[dl] = True
                tmp = torch.zeros_like(stop)
                tmp[chk] = slow | div
                stop |= tmp
            it[loc[stop]] = False
            go = ~stop
            gl, gg = loc[go], g[go]
            OldNrm[gl] = torch.clamp(nn[go], min=EPS)
            wn = w[gl] + zz[go]
            w[gl] = wn
            z[gl] = torch.einsum("jq,mqn->mjn", T, wn)
            done = self.FacConv[gg] * nn[go] <= FNewt
            it[gl[done]] = False
        ok = ~failed
        if not bool(ok.any()):
            return
        loc = ok.nonzero().squeeze(1)
        g = idx[loc]
        zg, Sg = z[loc], Scal[loc]
        # ------- Estrad (radaute_impl.hpp:1902-1986)
        temp = torch.einsum("mqn,mq->mn", zg, Dd[None, :] * (1.0 / self.h[g])[:, None])   # z * Dd/h
        rhs = self.f0[g] + temp
        ev = self._lusolve(g, 0, torch.complex(rhs, torch.zeros_like(rhs))).real
        err = torch.clamp((ev / Sg).norm(dim=1) / Ny ** 0.5, min=1e-10)
        retry = (err >= 1) & (self.First[g] | self.Reject[g])
        if bool(retry.any()):
            rg = g[retry]
            fv = self.ode(self.t[rg], self.y[rg] + err[retry][:, None])
            self.stats[rg, ST["Fcn"]] += 1
            rhs2 = fv + temp[retry]
            ev2 = self._lusolve(rg, 0, torch.complex(rhs2, torch.zeros_like(rhs2))).real
            err[retry] = torch.clamp((ev2 / Sg[retry]).norm(dim=1) / Ny ** 0.5, min=1e-10)
        badn = ~torch.isfinite(err)
        if bool(badn.any()):
            self.res[g[badn]] = 2
            self.active[g[badn]] = False
            keep = ~badn
            loc, g, zg, err = loc[keep], g[keep], zg[keep], err[keep]
            if g.numel() == 0:
                return
        newt = self.Newt[g].double()
        fac = torch.clamp((2.0 * Nit + 1.0) / (2.0 * Nit + newt), max=op.Safe)
        FacL, FacR = 1.0 / op.FacL, 1.0 / op.FacR
        quot = torch.clamp(err ** (1.0 / s + 1.0) / fac, FacR, FacL)
        hnew = self.h[g] / quot
        acc = err < 1
        # ---- accepted
        if bool(acc.any()):
            a, za, ea = g[acc], zg[acc], err[acc]
            self.First[a] = False
            self.stats[a, ST["Acc"]] += 1
            hn = hnew[acc]
            qa = quot[acc]
            if op.Gustafsson:
                gm = ~self.ChangeFlag[a]
                lim = torch.ones_like(gm) if (getattr(op, "RadauTeQuirks", 1) & 1) else gm   # quirk Q8
                gg2 = lim & (self.stats[a, ST["Acc"]] > 1)
                if bool(gg2.any()):
                    ag = a[gg2]
                    fg = (self.hacc[ag] / self.h[ag]) * (ea[gg2] ** 2 / self.erracc[ag]) ** (1.0 / (s + 1.0)) / op.Safe
                    fg = torch.clamp(fg, FacR, FacL)
                    qn = torch.maximum(qa[gg2], fg)
                    qa[gg2] = qn
                    hn[gg2] = self.h[ag] / qn
                self.hacc[a[gm]] = self.h[a[gm]]
                self.erracc[a[gm]] = torch.clamp(ea[gm], min=1.0e-2)
            self.h_old[a] = self.h[a]
            self.t[a] = self.t[a] + self.h[a]
            self.y[a] = self.y[a] + za[:, s - 1]
            cont = torch.zeros(a.numel(), 7, Ny, dtype=torch.float64)
            cont[:, s - 1] = za[:, 0] / C[0]
            for q in range(1, s):
                cont[:, q - 1] = (za[:, s - q - 1] - za[:, s - q]) * (1.0 / (C[s - q - 1] - C[s - q]))
            for jj in range(2, s + 1):
                for k in range(s, jj - 1, -1):
                    Fact = 1.0 / (-C[jj - 1]) if s - k == 0 else 1.0 / (C[s - k - 1] - C[s - k + jj - 1])
                    cont[:, k - 1] = (cont[:, k - 1] - cont[:, k - 2]) * Fact
            self.cont[a] = cont
            self.NeedNewJac[a] = True
            fin = self.Last[a]
            if bool(fin.any()):
                af = a[fin]
                self.h[af] = self.hopt[af]
                self.stats[af, ST["Rej"]] += 1
                self.active[af] = False
            go = ~fin
            if bool(go.any()):
                b, hb = a[go], hn[go]
                fv = self.ode(self.t[b], self.y[b])
                self.f0[b] = fv
                self.stats[b, ST["Fcn"]] += 1
                nanrow = torch.isnan(fv).any(1)
                if bool(nanrow.any()):
                    self.res[b[nanrow]] = 2
                    self.active[b[nanrow]] = False
                hb = self.sgn[b] * torch.minimum(hb.abs(), self.hmaxn[b])
                self.hopt[b] = self.sgn[b] * torch.minimum(self.h[b].abs(), hb.abs())
                self.Reject[b] = False
                lastm = (self.t[b] + hb / op.Quot1 - self.tf[b]) * self.sgn[b] >= 0
                bl = b[lastm]
                self.h[bl] = self.tf[bl] - self.t[bl]
                self.Last[bl] = True
                nl = ~lastm
                bn, hbn = b[nl], hb[nl]
                qt = hbn / self.h[bn]
                keepm = (self.Theta[bn] <= op.JacRecompute) & (qt >= op.Quot1) & (qt <= op.Quot2)
                bk = bn[keepm]
                self.Keep[bk] = True
                self.NeedNewJac[bk] = False
                self.NeedNewLU[bk] = False
                bnk = bn[~keepm]
                self.h[bnk] = hbn[~keepm]
                rest = torch.cat([bl, bnk])
                self.NeedNewLU[rest] = True
                self.NeedNewJac[rest[self.Theta[rest] <= op.JacRecompute]] = False
        # ---- rejected
        rej = ~acc
        if bool(rej.any()):
            r = g[rej]
            self.Reject[r] = True
            self.Last[r] = False
            first = self.First[r]
            self.h[r] = torch.where(first, self.h[r] / 10.0, hnew[rej])
            self.stats[r[self.stats[r, ST["Acc"]] >= 1], ST["Rej"]] += 1
            self.NeedNewLU[r] = True

    # ------------------------------------------------------------------ driver
    def solve(self):
        op = self.op
        Variab = op.MaxNbrStg != op.MinNbrStg
        while True:
            A = self.active & (self.stats[:, ST["Step"]] <= op.MaxNbrStep) & (self.sgn * self.t < self.sgn * self.tf)
            capped = self.active & ~A & (self.sgn * self.t < self.sgn * self.tf)
            self.res[capped & (self.res == 0)] = 3
            self.active &= A
            if not bool(A.any()):
                break
            self.count += 1
            idx = A.nonzero().squeeze(1)
            self.stats[idx, ST["Step"]] += 1
            self.FacConv[idx] = torch.clamp(self.FacConv[idx], min=EPS) ** 0.8
            j = idx[self.NeedNewJac[idx]]
            if j.numel():
                self.Jac[j] = self.jac(self.t[j], self.y[j])
                self.stats[j, ST["Jac"]] += 1
                self.NeedNewJac[j] = False
                self.NeedNewLU[j] = True
            if Variab:
                v = idx[~self.Keep[idx]]
                if v.numel():
                    self.ChangeNbr[v] += 1
                    ns = self.NbrStg[v].clone()
                    cur = self.NbrStg[v]
                    hq = self.h[v] / self.h_old[v]
                    self.Thetat[v] = torch.clamp(torch.maximum(self.Theta[v], self.Thetat[v] * 0.5), max=10.0)
                    up = (self.Newt[v] > 1) & (self.Thetat[v] <= op.Vitu) & (hq < op.hhou) & (hq > op.hhod)
                    ns[up] = torch.clamp(cur[up] + 2, max=op.MaxNbrStg)
                    dn = (self.Thetat[v] >= op.Vitd) | self.UnExpStepRej[v]
                    ns[dn] = torch.clamp(cur[dn] - 2, min=op.MinNbrStg)
                    dn2 = (self.ChangeNbr[v] >= 1) & self.UnExpNewtRej[v]
                    ns[dn2] = torch.clamp(cur[dn2] - 2, min=op.MinNbrStg)
                    lim = self.ChangeNbr[v] <= 10
                    ns[lim] = torch.minimum(cur[lim], ns[lim])
                    ch = ns != cur
                    self.ChangeFlag[v] = ch
                    self.UnExpNewtRej[v] = False
                    self.UnExpStepRej[v] = False
                    vc = v[ch]
                    self.NbrStg[vc] = ns[ch]
                    self.ChangeNbr[vc] = 1
                    self.NeedNewLU[vc] = True
            for s in (1, 3, 5, 7):
                g = idx[self.NbrStg[idx] == s]
                if g.numel():
                    self._group_pass(g, s)
        self.nstg = self.NbrStg
        return self.res.double()


def hamiltonian_callbacks(H, W=None, control=None, sign=+1):
    """OdeFcn / JacFn pair of the reference's Hamiltonian example
    (examples/cpp/radaute_hamiltonian_example.cpp:44-62): evalDyns / evalJac by autograd."""
    from . import ham

    def _u(y):
        if control is None:
            return None
        N = y.shape[1] // 2
        return control(y[:, N:], y[:, :N], W).detach()

    def ode(t, y):
        return ham.eval_dyns(y, W, H, _u(y), costate_sign=sign)

    def jac(t, y):
        return ham.eval_jac(y, W, H, _u(y), costate_sign=sign)
    return ode, jac
