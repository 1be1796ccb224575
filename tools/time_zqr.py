"""ComplexF64 blocked QR timing: python tools/time_zqr.py n [n ...]   (GLA_ZGEMM_FMA=1 selects the FMA contraction)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
g = ge.load()
sizes = [int(x) for x in sys.argv[1:]] or [2048, 4096]
st = torch.cuda.current_stream().cuda_stream
for n in sizes:
    src = torch.randn((n, n), device="cuda", dtype=torch.complex128)
    dA = src.clone()
    dtau = torch.zeros(n, device="cuda", dtype=torch.complex128)
    ts = []
    for it in range(3):
        dA.copy_(src)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, st, np.complex128)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"zqr n={n}: {ms:.2f} ms  {4*4/3*n**3/ms/1e9:.2f} real TFLOP/s  (all: {[round(t,1) for t in ts]})", flush=True)
    if n <= 8192:
        F = dA.t()                      # storage is column-major: dA[j, i] = F[i, j]
        R = torch.triu(F)
        A0 = src.t()
        G = A0.conj().t() @ A0
        err = (R.conj().t() @ R - G).abs().max() / G.abs().max()
        print("   gram err", err.item(), flush=True)
        E = (R.conj().t() @ R - G).abs()
        blk = max(n // 8, 1)
        print("   per column block:", [f"{E[:, c:c + blk].max().item() / G.abs().max().item():.1e}" for c in range(0, n, blk)], flush=True)
