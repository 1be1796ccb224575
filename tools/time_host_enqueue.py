"""Host enqueue time vs GPU time of the launch-chain-bound paths."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
g = ge.load()
st = torch.cuda.current_stream().cuda_stream
n = 4096
X = torch.randn((n, n), device="cuda", dtype=torch.float64)
S = X.t() @ X + n * torch.eye(n, device="cuda", dtype=torch.float64)
info = torch.zeros(1, device="cuda", dtype=torch.int32)
for it in range(4):
    dA = S.clone(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    g.chol_recursive_dev(dA.data_ptr(), n, n, info.data_ptr(), 1, st)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"chol n={n}: host enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms", flush=True)
for n in (1024, 4096):
    src = torch.randn((n, n), device="cuda", dtype=torch.float64); tau = torch.zeros(n, device="cuda", dtype=torch.float64)
    for it in range(3):
        dA = src.clone(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.qr_blocked_dev(dA.data_ptr(), n, n, n, tau.data_ptr(), 0, st)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"qr n={n}: host enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms", flush=True)
