"""Device time of the two-sided reductions (persistent cooperative kernels) at a few sizes; optional oracle timing.
usage: python tools/time_twosided.py [n ...]   (CPU=1 also times the oracle on the host cores)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

gla = ge.load()
sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096]
st = torch.cuda.current_stream().cuda_stream
for n in sizes:
    g = torch.Generator(device="cuda").manual_seed(n)
    A0 = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
    S0 = A0 + A0.T
    t1 = torch.zeros(n, dtype=torch.float64, device="cuda")
    t2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    out = {}
    for name, src, call in (
        ("bidiag", A0, lambda W: gla.bidiagonalize_dev(W.data_ptr(), n, n, n, t1.data_ptr(), t2.data_ptr(), st)),
        ("hessenberg", A0, lambda W: gla.hessenberg_dev(W.data_ptr(), n, n, t1.data_ptr(), st)),
        ("symtriL", S0, lambda W: gla.symtri_dev(W.data_ptr(), n, n, "L", t1.data_ptr(), st)),
        ("symtriU", S0, lambda W: gla.symtri_dev(W.data_ptr(), n, n, "U", t1.data_ptr(), st)),
    ):
        ts = []
        for rep in range(3):
            W = src.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call(W)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name] = min(ts[1:])
    flops = {"bidiag": 8 / 3 * n ** 3, "hessenberg": 10 / 3 * n ** 3, "symtriL": 4 / 3 * n ** 3, "symtriU": 4 / 3 * n ** 3}
    print(f"n={n}: " + "  ".join(f"{k} {v:8.2f} ms ({flops[k] / v / 1e9:6.2f} TFLOP/s, {v * 1e3 / n:5.1f} us/step)" for k, v in out.items()),
          flush=True)
    if os.environ.get("CPU") == "1" and n <= 2048:
        from oracle import oracle as O
        A = A0.cpu().numpy().T.copy(order="F")
        for name, f in (("bidiag", lambda: O.bidiagonalize(A.copy(order="F"))), ("hessenberg", lambda: O.hessenberg(A.copy(order="F"))),
                        ("symtriL", lambda: O.symtri(np.asfortranarray(A + A.T), "L"))):
            t0 = time.perf_counter()
            f()
            print(f"   oracle {name}: {(time.perf_counter() - t0) * 1e3:9.1f} ms on {O.max_threads()} threads", flush=True)
