"""Determinism / race stress: the blocked QR is deterministic by construction, so repeated factorisations of the
same input must be BITWISE equal.  python tools/stress_qr.py dtype n reps [noise]   (dtype: d | z | s; s = Float32 on tcgen05)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
g = ge.load()
kind = sys.argv[1]; n = int(sys.argv[2]); reps = int(sys.argv[3])
noise_mode = sys.argv[4] if len(sys.argv) > 4 else "none"     # none | normal | high: unrelated kernels on another stream
dt = {"z": torch.complex128, "s": torch.float32}.get(kind, torch.float64)
npdt = {"z": np.complex128, "s": np.float32}.get(kind, np.float64)
st = torch.cuda.current_stream().cuda_stream
src = torch.randn((n, n), device="cuda", dtype=dt)
dtau = torch.zeros(n, device="cuda", dtype=dt)
ref = None
bad = 0
noise = None if noise_mode == "none" else torch.cuda.Stream(priority=-1 if noise_mode == "high" else 0)
nz = [torch.zeros(1 << 22, device="cuda") for _ in range(4)]
big = torch.zeros((2048, 2048), device="cuda")
for it in range(reps):
    dA = src.clone()
    torch.cuda.synchronize()
    g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, st, npdt)
    if noise is not None:
        with torch.cuda.stream(noise):
            for k in range(300):
                nz[k & 3].add_(1.0)
                if k % 50 == 0:
                    torch.mm(big, big)
    torch.cuda.synchronize()
    # cheap per-rep correctness probe: R^H R x = A^H A x for a random x (storage is column-major: dA[j, i] = F[i, j])
    xv = torch.randn(n, device="cuda", dtype=dt, generator=None)
    Rm = torch.triu(dA.t())
    A0m = src.t()
    y1 = Rm.conj().t() @ (Rm @ xv)
    y2 = A0m.conj().t() @ (A0m @ xv)
    perr = ((y1 - y2).abs().max() / y2.abs().max()).item()
    if perr > 1e-10:
        print(f"  rep {it}: PROBE residual {perr:.2e}", flush=True)
    del Rm, y1, y2
    if ref is None:
        ref = dA
        A0 = src.t(); R = torch.triu(dA.t())
        G = A0.conj().t() @ A0
        print(f"{kind} n={n} first run gram err {((R.conj().t() @ R - G).abs().max() / G.abs().max()).item():.2e}", flush=True)
    else:
        diff = (torch.view_as_real(dA) if kind == "z" else dA) != (torch.view_as_real(ref) if kind == "z" else ref)
        nd = int(diff.sum().item())
        if nd:
            bad += 1
            idx = diff.nonzero()[0].tolist()
            mx = (dA - ref).abs().max().item()
            print(f"  rep {it}: {nd} entries differ, first at {idx}, max abs diff {mx:.3e}", flush=True)
print(f"{kind} n={n}: {bad}/{reps - 1} repeats differed from the first run", flush=True)
