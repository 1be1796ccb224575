"""CPU suite, part 2: the C-ABI library loads without a GPU and exports every symbol include/gla_cuda.h
declares; the host mirror maps argument errors like the reference; the product never imports oracle/."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gla_cuda.h")


def _declared():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"GLA_API\s+[\w\s\*]+?\b(gla_\w+)\s*\(", txt)))


def test_library_builds_loads_and_exports_every_declared_symbol(gla):
    import __graft_entry__ as ge
    if not os.path.exists(gla.glacuda.LIB_PATH):
        ge.build()
    lib = ctypes.CDLL(gla.glacuda.LIB_PATH)
    names = _declared()
    assert len(names) >= 38
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gla_cuda.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", gla.glacuda.LIB_PATH], text=True)
    exported = set(re.findall(r" T (gla_\w+)", out))
    assert exported == set(names), exported ^ set(names)
    assert lib.gla_version() >= 1


def test_library_targets_sm_100a_only(gla):
    out = subprocess.run(["cuobjdump", "-lelf", gla.glacuda.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_mirror_argument_errors_without_gpu(gla):
    with pytest.raises(gla.DimensionMismatch):
        gla.reflectorApply_(np.zeros((5, 5), order="F"), np.zeros(4), 1.0)          # test/qr.jl:29-33
    with pytest.raises(gla.DimensionMismatch):
        gla.cholRecursive_(np.zeros((3, 4), order="F"))                              # checksquare
    with pytest.raises(gla.ArgumentError):
        gla.cholRecursive_(np.zeros((3, 3), order="F"), "U")                         # only Val{:L} has a method
    with pytest.raises(gla.ArgumentError):
        gla.QR2(np.zeros((5, 10), order="F"), np.zeros(5)).R                         # test/qr.jl:34
    with pytest.raises(TypeError):
        gla.qrBlocked_(np.zeros((4, 4), dtype=np.float16, order="F"))                # stays on the reference path
    with pytest.raises(gla.ArgumentError):
        gla.qrBlocked_(np.zeros((4, 6))[:, ::2])                                     # not unit row stride
    # empty problems return before touching the device
    q = gla.qrBlocked_(np.zeros((0, 5), order="F"))
    assert q.tau.size == 0
    gla.qr_batched_(np.zeros((0, 32, 32)))


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genericlinearalgebra.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), f
