"""Float32 qrBlocked! timings (3xTF32 tensor-pipe contraction vs GLA_SGEMM_FMA=1) and a Gram check."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
sizes = [int(x) for x in sys.argv[1:]] or [2048, 8192, 16384]
st = torch.cuda.current_stream().cuda_stream
for n in sizes:
    src = torch.randn((n, n), device="cuda", dtype=torch.float32)
    dA = src.clone()
    dtau = torch.zeros(n, device="cuda", dtype=torch.float32)
    ts = []
    for it in range(4):
        dA.copy_(src)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, st, np.float32)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    x = torch.randn(n, device="cuda", dtype=torch.float64)
    R = torch.triu(dA.t()).double()
    y1 = R.t() @ (R @ x)
    s64 = src.double()
    y2 = s64 @ (s64.t() @ x)
    print(f"f32 n={n}: {ms:.2f} ms  {4/3*n**3/ms/1e9:.2f} TFLOP/s  gram probe {((y1-y2).abs().max()/y2.abs().max()).item():.2e}", flush=True)
