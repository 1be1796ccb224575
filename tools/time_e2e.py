"""End-to-end (host pointers, pinned) batched QR timing: python tools/time_e2e.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
g = ge.load()
batch = 1 << 20
hA = torch.randn((batch, 32, 32), dtype=torch.float64).pin_memory()
ht = torch.empty((batch, 32), dtype=torch.float64).pin_memory()
g.qr_batched_ptr(hA.data_ptr(), 32, 32, 65536, ht.data_ptr())
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    g.qr_batched_ptr(hA.data_ptr(), 32, 32, batch, ht.data_ptr())
    ts.append(time.perf_counter() - t0)
print(f"chunk_mb={os.environ.get('GLA_BATCH_CHUNK_MB','64')} streams={os.environ.get('GLA_BATCH_STREAMS','3')}: "
      f"best {min(ts)*1e3:.1f} ms -> {batch/min(ts)/1e6:.2f} M matrices/s, {(8192*2+256)*batch/min(ts)/1e9:.1f} GB/s PCIe total  (all {[round(t*1e3) for t in ts]})", flush=True)
# raw PCIe reference: plain pinned copies of the same volume, both directions concurrently
d = torch.empty((batch, 32, 32), dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(hA, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
with torch.cuda.stream(s2): hA.copy_(d, non_blocking=True)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"raw H2D {8.59/(t1-t0):.1f} GB/s, raw D2H {8.59/(t2-t1):.1f} GB/s", flush=True)
