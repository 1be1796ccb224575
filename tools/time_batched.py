import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import __graft_entry__ as ge
g = ge.load()
for dt, npdt in ((torch.float64, np.float64), (torch.float32, np.float32)):
    batch = 1 << 20 if dt == torch.float64 else 1 << 20
    dA = torch.randn((batch, 32, 32), device="cuda", dtype=dt)
    src = dA.clone()
    dtau = torch.empty((batch, 32), device="cuda", dtype=dt)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        dA.copy_(src)
        g.qr_batched_dev(dA.data_ptr(), 32, 32, batch, dtau.data_ptr(), st, npdt)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        dA.copy_(src)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        g.qr_batched_dev(dA.data_ptr(), 32, 32, batch, dtau.data_ptr(), st, npdt)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    el = dA.element_size()
    byt = batch * (2 * 1024 * el + 32 * el)
    print(f"{dt}: best {ms:.3f} ms median {sorted(ts)[5]:.3f} -> {batch/ms/1e3:.1f} M mat/s, {byt/ms/1e6:.0f} GB/s, {batch*43691/ms/1e9:.2f} TFLOP/s")
