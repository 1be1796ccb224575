"""Time one panel factorisation (m x 64 Float64 -> a single qr_panel_kernel launch) with CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
g = ge.load()
st = torch.cuda.current_stream().cuda_stream
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for m in (64, 128, 192, 1024, 4096, 8192, 16384):
    src = torch.randn((nb, m), device="cuda", dtype=torch.float64)
    dA = src.clone()
    dtau = torch.zeros(nb, device="cuda", dtype=torch.float64)
    ts = []
    for it in range(12):
        dA.copy_(src)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        g.qr_blocked_dev(dA.data_ptr(), m, nb, m, dtau.data_ptr(), 0, st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    print(f"m={m:6d} nb={nb}: min {ts[0]:7.1f} us  median {ts[len(ts)//2]:7.1f} us", flush=True)
