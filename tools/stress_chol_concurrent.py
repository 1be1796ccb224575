"""Does a TMA/DMMA kernel chain stay bitwise reproducible while OTHER streams run small kernels concurrently?
python tools/stress_chol_concurrent.py n reps mode     mode: none | normal | high   (priority of the noise stream)
The Cholesky driver is single-stream and uses the same gemm_tn DMMA kernels as the blocked QR."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
g = ge.load()
n = int(sys.argv[1]); reps = int(sys.argv[2]); mode = sys.argv[3]
X = torch.randn((n, n), device="cuda", dtype=torch.float64)
S = X.t() @ X + n * torch.eye(n, device="cuda", dtype=torch.float64)
info = torch.zeros(1, device="cuda", dtype=torch.int32)
main = torch.cuda.Stream()
noise = None if mode == "none" else torch.cuda.Stream(priority=-1 if mode == "high" else 0)
nz = [torch.zeros(1 << 22, device="cuda") for _ in range(4)]
big = torch.zeros((4096, 4096), device="cuda")
ref = None
bad = 0
torch.cuda.synchronize()
for it in range(reps):
    dA = S.clone()
    torch.cuda.synchronize()
    with torch.cuda.stream(main):
        g.chol_recursive_dev(dA.data_ptr(), n, n, info.data_ptr(), 1, main.cuda_stream)
    if noise is not None:
        with torch.cuda.stream(noise):
            for k in range(300):
                nz[k & 3].add_(1.0)
                if k % 50 == 0:
                    torch.mm(big, big)        # something that wants every SM for a moment
    torch.cuda.synchronize()
    if ref is None:
        ref = dA
        L = torch.tril(dA.t())
        print(f"chol n={n} mode={mode}: first run resid {((L @ L.t() - S).norm() / S.norm()).item():.2e}", flush=True)
    else:
        nd = int((torch.tril(dA.t()) != torch.tril(ref.t())).sum().item())
        if nd:
            bad += 1
            print(f"  rep {it}: {nd} entries differ, max abs {(torch.tril(dA.t()) - torch.tril(ref.t())).abs().max().item():.3e}", flush=True)
print(f"chol n={n} mode={mode}: {bad}/{reps - 1} repeats differed", flush=True)
