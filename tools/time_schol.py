"""Float32 Cholesky timing: python tools/time_schol.py [n ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
st = torch.cuda.current_stream().cuda_stream
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]:
    X = torch.randn((n, n), device="cuda", dtype=torch.float32)
    S = X.t() @ X + n * torch.eye(n, device="cuda", dtype=torch.float32)
    info = torch.zeros(1, device="cuda", dtype=torch.int32)
    ts = []
    for it in range(4):
        dS = S.clone()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g.chol_recursive_dev(dS.data_ptr(), n, n, info.data_ptr(), 1, st, np.float32); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    L = torch.tril(dS.t()).double()
    res = ((L @ L.t() - S.double()).norm() / S.double().norm()).item()
    print(f"chol f32 n={n}: {min(ts[1:]):.2f} ms  {n**3/3/min(ts[1:])/1e9:.2f} TFLOP/s  residual {res:.2e} info {int(info.item())}", flush=True)
