"""Float32 contraction probe: rankUpdate! (C += alpha A A^H, lower) drives gemm_tn<float> directly; compared with a
Float64 product.  usage: python tools/probe_umma.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

g = ge.load()
rng = np.random.default_rng(0)
for (n, k) in [(128, 32), (128, 64), (256, 96), (300, 200), (1000, 520), (2048, 4096)]:
    A = np.asfortranarray(rng.standard_normal((n, k)).astype(np.float32))
    C0 = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32))
    C = C0.copy(order="F")
    t0 = time.perf_counter()
    g.rankUpdate_(C, A, -1.0)
    dt = time.perf_counter() - t0
    ref = C0.astype(np.float64) - A.astype(np.float64) @ A.astype(np.float64).T
    low = np.tril_indices(n)
    err = np.max(np.abs(C[low] - ref[low])) / np.max(np.abs(ref))
    up = np.array_equal(np.triu(C, 1), np.triu(C0, 1))
    print(f"n={n} k={k}: max rel err {err:.3e}, upper untouched {up}, {dt * 1e3:.1f} ms", flush=True)
