"""host-pointer gla_dtsqr on a pinned 8,388,608 x 64 matrix (4.3 GB): python tools/time_e2e_tsqr.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
m, n = 1 << 23, 64
h = torch.empty((n, m), dtype=torch.float64, pin_memory=True)      # column-major m x n
for c in range(0, n, 8):
    h[c:c + 8].copy_(torch.randn((8, m), device="cuda", dtype=torch.float64))
A = h.numpy().T
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    R = g.tsqr_R(A)
    ts.append((time.perf_counter() - t0) * 1e3)
G = torch.zeros((n, n), device="cuda", dtype=torch.float64)
for r0 in range(0, m, 1 << 21):
    blk = h[:, r0:r0 + (1 << 21)].cuda()
    G += blk @ blk.t()
Gh = G.cpu().numpy()
err = np.abs(R.T @ R - Gh).max() / np.abs(Gh).max()
print(f"e2e tsqr {m}x{n}: {min(ts[1:]):.1f} ms (all {[round(t, 1) for t in ts]}) = {m * n * 8 / min(ts[1:]) / 1e6:.1f} GB/s of host data, gram err {err:.1e}", flush=True)
