"""One blocked-QR factorisation (after `warm` untimed ones) for ncu launch lists: python tools/prof_qr.py n [warm] [dtype]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
g = ge.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
st = torch.cuda.current_stream().cuda_stream
src = torch.randn((n, n), device="cuda", dtype=torch.float64)
dA = src.clone()
dtau = torch.zeros(n, device="cuda", dtype=torch.float64)
for it in range(warm + 1):
    dA.copy_(src)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, st)
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} iter {it}: {e0.elapsed_time(e1):.2f} ms", flush=True)
