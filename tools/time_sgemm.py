"""Float32 contraction throughput on device-resident data: times the wide block application (ormqr_blocked_dev with 384
reflectors = W = V^H A and A -= V Z, the two big products of the blocked QR's far update, plus the panel preparation).
usage: python tools/time_sgemm.py [nA ...]   (REPS=1 for ncu launch lists, GLA_SGEMM_MMASYNC=1 for the mma.sync kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
st = torch.cuda.current_stream().cuda_stream
m, k = 16384, 384
sizes = [int(a) for a in sys.argv[1:]] or [2048, 8192, 16384]
for nA in sizes:
    F = torch.randn((k, m), device="cuda", dtype=torch.float32)        # column-major m x k factors
    tau = torch.rand(k, device="cuda", dtype=torch.float32) + 0.5
    A = torch.randn((nA, m), device="cuda", dtype=torch.float32)       # column-major m x nA
    ts = []
    for it in range(int(os.environ.get('REPS', '4'))):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        g.ormqr_blocked_dev(F.data_ptr(), m, k, m, tau.data_ptr(), A.data_ptr(), m, nA, m, True, st, np.float32)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:]) if len(ts) > 1 else ts[0]
    fl = 4.0 * m * k * nA   # W = V^H A and A -= V Z
    print(f"ormqr f32 m={m} k={k} nA={nA}: {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s (two K/M=384 products + panel prep)", flush=True)
