"""host-pointer gla_dpotrf_recursive_L on a pinned matrix: python tools/time_e2e_chol.py [n ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
    X = torch.randn((n, n), device="cuda", dtype=torch.float64)
    S = X.t() @ X + n * torch.eye(n, device="cuda", dtype=torch.float64)
    hS = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
    ts = []
    for _ in range(3):
        hS.copy_(S)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.cholRecursive_(hS.numpy().T)      # symmetric input: the transposed view is the column-major matrix
        ts.append((time.perf_counter() - t0) * 1e3)
    got = hS.numpy().T
    L = np.tril(got)
    Sn = S.cpu().numpy()
    res = np.linalg.norm(L @ L.T - Sn) / np.linalg.norm(Sn)
    up = np.array_equal(np.triu(got, 1), np.triu(Sn.T, 1))
    print(f"e2e chol f64 n={n}: {min(ts[1:]):.2f} ms (all {[round(t, 2) for t in ts]}), device part {g.glacuda.lib().gla_last_device_ms():.2f} ms, "
          f"residual {res:.1e}, strict upper triangle untouched {up}", flush=True)
