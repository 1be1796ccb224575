"""TSQR local kernel timing: python tools/time_tsqr.py [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
g = ge.load()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
st = torch.cuda.current_stream().cuda_stream
for n in (64, 32):
    A = torch.randn((n, m), device="cuda", dtype=torch.float64)
    R = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    ts = []
    for it in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g.tsqr_local_dev(A.data_ptr(), m, n, m, R.data_ptr(), n, st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    Ru = torch.triu(R.t()); G = A @ A.t()
    err = ((Ru.t() @ Ru - G).abs().amax() / G.abs().amax()).item()
    print(f"tsqr {m}x{n}: {ms:.3f} ms  {2*m*n*n/ms/1e9:.2f} TFLOP/s  {m*n*8/ms/1e6:.0f} GB/s  gram err {err:.2e}  (all {[round(t,2) for t in ts]})", flush=True)
