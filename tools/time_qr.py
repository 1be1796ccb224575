import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
sizes = [int(x) for x in sys.argv[1:]] or [1024, 4096, 8192, 16384]
st = torch.cuda.current_stream().cuda_stream
for n in sizes:
    src = torch.randn((n, n), device="cuda", dtype=torch.float64)
    dA = src.clone()
    dtau = torch.zeros(n, device="cuda", dtype=torch.float64)
    ts = []
    for it in range(4):
        dA.copy_(src)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"n={n}: {ms:.2f} ms  {4/3*n**3/ms/1e9:.2f} TFLOP/s  (all: {[round(t,1) for t in ts]})", flush=True)
    # residual check via R^T R = A^T A on a slice (storage is column-major: dA[j, i] = F[i, j])
    if n <= 8192:
        F = dA.t()
        R = torch.triu(F)
        A0 = src.t()
        err = (R.t() @ R - A0.t() @ A0).abs().max() / (A0.t() @ A0).abs().max()
        print("   gram err", err.item())
# DGEMM reference (cuBLAS via torch) for the FP64 roofline denominator
a = torch.randn((8192, 8192), device="cuda", dtype=torch.float64); b = torch.randn((8192, 8192), device="cuda", dtype=torch.float64)
for _ in range(2): a @ b
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
print(f"cuBLAS DGEMM 8192^3: {2*8192**3/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s")
