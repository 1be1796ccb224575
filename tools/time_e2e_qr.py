"""host-pointer gla_dgeqr_blocked on a pinned matrix (H2D + factorisation + D2H): python tools/time_e2e_qr.py [n ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
for n in [int(a) for a in sys.argv[1:]] or [4096, 16384]:
    hA = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
    htau = torch.empty(n, dtype=torch.float64, pin_memory=True)
    src = torch.randn((n, n), device="cuda", dtype=torch.float64)
    ts = []
    for _ in range(3):
        hA.copy_(src)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.qr_blocked_ptr(hA.data_ptr(), n, n, n, htau.data_ptr(), 0, np.float64)
        ts.append((time.perf_counter() - t0) * 1e3)
    # check against a device-resident factorisation of the same matrix
    dA = src.clone(); dtau = torch.zeros(n, device="cuda", dtype=torch.float64)
    g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, torch.cuda.current_stream().cuda_stream, np.float64)
    torch.cuda.synchronize()
    same = torch.equal(hA.cuda(), dA) and torch.equal(htau.cuda(), dtau)
    print(f"e2e qr f64 n={n}: {min(ts[1:]):.1f} ms (all {[round(t, 1) for t in ts]}), device-resident last {g.glacuda.lib().gla_last_device_ms():.1f} ms, "
          f"host result bitwise equal to the device-resident one: {same}", flush=True)
