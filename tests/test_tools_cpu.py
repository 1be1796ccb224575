"""The variant-sweep checker (tools/sweep_check.py) on synthetic dumps: a dump that holds the oracle's own result passes,
a perturbed one is flagged.  Keeps the torch-free GPU harness (tools/sweep_batched.cu) and its checker in step."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_checker():
    spec = importlib.util.spec_from_file_location("sweep_check", os.path.join(ROOT, "tools", "sweep_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _dump(path, oracle, nd, perturb=0.0):
    rng = np.random.default_rng(5)
    with open(path, "wb") as f:
        for _part in range(2):
            A = rng.standard_normal((nd, 32, 32))
            fac, tau = oracle.qr_batched(A)
            fac = np.array(fac)
            fac[nd // 2, 3, 7] += perturb
            np.ascontiguousarray(np.transpose(A, (0, 2, 1))).tofile(f)       # column-major per matrix
            np.ascontiguousarray(np.transpose(fac, (0, 2, 1))).tofile(f)
            np.ascontiguousarray(tau).tofile(f)


def test_sweep_check_accepts_oracle_and_flags_perturbation(tmp_path, oracle):
    chk = _load_checker()
    good = str(tmp_path / "good.bin")
    bad = str(tmp_path / "bad.bin")
    _dump(good, oracle, 6)
    _dump(bad, oracle, 6, perturb=1e-6)
    n, worst, wt, ok = chk.check(good)
    assert n == 12 and ok and worst == 0.0 and wt == 0.0
    assert not chk.check(bad)[3]
    assert chk.main([good]) == 0 and chk.main([good, bad]) == 1
    with open(str(tmp_path / "junk.bin"), "wb") as f:
        f.write(b"\0" * 24)
    with pytest.raises(ValueError):
        chk.check(str(tmp_path / "junk.bin"))
