"""ComplexF64 Cholesky timing: python tools/time_zchol.py [n ...]   (GLA_CHOL_Z_RECURSIVE=1: round-1 recursion)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
st = torch.cuda.current_stream().cuda_stream
for n in [int(a) for a in sys.argv[1:]] or [2048, 4096]:
    X = torch.randn((n, n), device="cuda", dtype=torch.complex128)
    S = X.conj().t() @ X + n * torch.eye(n, device="cuda", dtype=torch.complex128)
    S = S.t().contiguous()   # column-major storage of S (Hermitian: S^T = conj(S))
    info = torch.zeros(1, device="cuda", dtype=torch.int32)
    ts = []
    for it in range(4):
        dS = S.clone()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g.chol_recursive_dev(dS.data_ptr(), n, n, info.data_ptr(), 1, st, np.complex128); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    L = torch.tril(dS.t())                      # dS[j, i] = A[i, j]
    Sref = S.t()
    res = ((L @ L.conj().t() - Sref).norm() / Sref.norm()).item()
    print(f"chol c128 n={n}: {min(ts[1:]):.2f} ms  {4 * n**3 / 3 / min(ts[1:]) / 1e9:.2f} real TFLOP/s  residual {res:.2e} info {int(info.item())}", flush=True)
