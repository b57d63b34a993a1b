"""The C-ABI library loads without a GPU, exports every symbol include/janus_b200.h declares,
validates options like the reference constructor (raising instead of exit(1)) and refuses to
compute without CUDA (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

import janus_ode_b200 as jb
from janus_ode_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "janus_b200.h")).read()
    declared = set(re.findall(r"\b(jb200_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("jb200_opts")
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for s in declared:
        assert getattr(L, s) is not None
    assert b"sm_100a" in L.jb200_version()


def test_default_options_match_reference_defaults():
    o = _lib.jb200_opts()
    assert _lib.lib().jb200_opts_default(C.byref(o)) == 0
    # src/cpp/radaute.hpp:52-96
    assert (o.RelTol[0], o.AbsTol[0], o.InitialStep, o.JacRecompute) == (1e-3, 1e-6, 1e-2, 1e-3)
    assert (o.MaxNbrNewton, o.NbrStg, o.MinNbrStg, o.MaxNbrStg, o.MaxNbrStep) == (7, 3, 3, 7, 10000)
    assert (o.Safe, o.Quot1, o.Quot2, o.FacL, o.FacR) == (0.9, 1.0, 1.2, 0.2, 8.0)
    assert (o.Vitu, o.Vitd, o.hhou, o.hhod, o.Gustafsson, o.costate_sign) == (0.002, 0.8, 1.2, 0.8, 1, 1)
    d = jb.OptionsTe()
    for k in ("InitialStep", "JacRecompute", "MaxNbrNewton", "NbrStg", "MinNbrStg", "MaxNbrStg", "MaxNbrStep",
              "Safe", "Quot1", "Quot2", "FacL", "FacR", "Vitu", "Vitd", "hhou", "hhod"):
        assert getattr(d, k) == getattr(o, k), k


@pytest.mark.parametrize("kw,msg", [
    (dict(NbrStg=4), "NbrStg must be 1, 3, 5 or 7"), (dict(MaxNbrNewton=3), "MaxNbrNewton"),
    (dict(RelTol=1e-17), "Relative tolerance"), (dict(AbsTol=-1.0), "AbsTol"), (dict(JacRecompute=1.5), "JacRecompute"),
    (dict(NbrStg=3, MinNbrStg=5, MaxNbrStg=7), "between MinNbrStg and MaxNbrStg"), (dict(Safe=1.5), "Safe"),
    (dict(Quot1=1.5), "Quot1"), (dict(Quot2=0.5), "Quot2"), (dict(FacL=2.0), "FacL"), (dict(FacR=0.5), "FacR"),
    (dict(hhou=0.5), "hhou"), (dict(RelTol=[1e-3, 1e-3, 1e-3]), "length 1 or 4")])
def test_option_errors_raise_like_the_reference_exits(kw, msg):
    y0 = torch.zeros(2, 4, dtype=torch.float64)
    ts = torch.tensor([[0.0, 1.0]] * 2, dtype=torch.float64)
    dyn = jb.HamiltonianSystem(jb.VanDerPol(10.0))
    with pytest.raises(ValueError, match=msg):
        jb.RadauTe(dyn, None, ts, y0, jb.OptionsTe(**kw))


def test_argument_errors():
    dyn = jb.HamiltonianSystem(jb.VanDerPol(10.0))
    y0 = torch.zeros(2, 4, dtype=torch.float64)
    ts = torch.tensor([[0.0, 1.0]] * 2, dtype=torch.float64)
    with pytest.raises(TypeError):
        jb.RadauTe(lambda t, y, p: y, None, ts, y0)                     # python callables cannot run on the GPU
    with pytest.raises(ValueError, match="strictly monotonic"):
        jb.RadauTe(dyn, None, torch.zeros(2, 2, dtype=torch.float64), y0)
    with pytest.raises(ValueError):
        jb.RadauTe(dyn, None, ts, torch.zeros(2, 3, dtype=torch.float64))
    with pytest.raises(TypeError):
        jb.RadauTe(dyn, None, ts.float(), y0.float())
    with pytest.raises(NotImplementedError):
        jb.RadauTe(dyn, None, ts, y0, jb.OptionsTe(EventsFcn=lambda *a: None))
    with pytest.raises(ValueError):
        jb.HamiltonianSystem(jb.VanDerPol(1.0), costate_sign=0)
    with pytest.raises(ValueError):
        jb.VdpPmp(control="nope")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    dyn = jb.HamiltonianSystem(jb.VanDerPol(10.0))
    y0 = torch.zeros(2, 4, dtype=torch.float64)
    ts = torch.tensor([[0.0, 1.0]] * 2, dtype=torch.float64)
    with pytest.raises(jb.JanusB200Error, match="no CPU fallback"):
        jb.RadauTe(dyn, None, ts, y0)                                   # handle creation needs a device
    with pytest.raises(jb.JanusB200Error):
        dyn(None, y0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "janus_ode_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/c/" not in txt.replace(
                    "mirrored by oracle/c/radau_oracle.c", "").replace("oracle/c/radau_oracle.c", ""), f
