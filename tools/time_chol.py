import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
st = torch.cuda.current_stream().cuda_stream
for n in [int(x) for x in sys.argv[1:]] or [1024, 4096, 8192]:
    X = torch.randn((n, n), device="cuda", dtype=torch.float64)
    S = X.t() @ X + n * torch.eye(n, device="cuda", dtype=torch.float64)
    dA = S.clone(); info = torch.zeros(1, device="cuda", dtype=torch.int32)
    ts = []
    import time
    for it in range(5):
        dA.copy_(S)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); g.chol_recursive_dev(dA.data_ptr(), n, n, info.data_ptr(), 1, st); e1.record()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append(e0.elapsed_time(e1))
        print(f"   it {it}: events {ts[-1]:.2f} ms, host enqueue {1e3*(t1-t0):.2f} ms, wall {1e3*(t2-t0):.2f} ms", flush=True)
    ms = min(ts[1:])
    L = torch.tril(dA.t())
    err = (L @ L.t() - S).norm() / S.norm()
    print(f"chol n={n}: {ms:.3f} ms {n**3/3/ms/1e9:.2f} TFLOP/s resid {err.item():.2e} info {info.item()}", flush=True)
# TSQR
m, n = 1 << 23, 64
A = torch.randn((n, m), device="cuda", dtype=torch.float64)
R = torch.zeros((n, n), device="cuda", dtype=torch.float64)
ts = []
for it in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g.tsqr_local_dev(A.data_ptr(), m, n, m, R.data_ptr(), n, st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts[1:])
print(f"tsqr {m}x{n}: {ms:.3f} ms  {2*m*n*n/ms/1e9:.2f} TFLOP/s  {m*n*8/ms/1e6:.0f} GB/s")
