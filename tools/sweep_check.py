"""Compare the dumps written by tools/sweep_batched against the CPU oracle (test infrastructure).
usage: python tools/sweep_check.py gpurun_out/sweep_v*.bin"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle

for path in sys.argv[1:]:
    raw = np.fromfile(path, dtype=np.float64)
    nd = raw.size // (2 * (2 * 1024 + 32))
    worst = 0.0
    wt = 0.0
    off = 0
    for part in range(2):
        a = raw[off:off + nd * 1024].reshape(nd, 32, 32); off += nd * 1024      # (batch, col, row)
        f = raw[off:off + nd * 1024].reshape(nd, 32, 32); off += nd * 1024
        t = raw[off:off + nd * 32].reshape(nd, 32); off += nd * 32
        rf, rt = oracle.qr_batched(np.transpose(a, (0, 2, 1)))
        rf = np.transpose(rf, (0, 2, 1))
        scale = np.max(np.abs(rf), axis=(1, 2), keepdims=True)
        worst = max(worst, float(np.max(np.abs(f - rf) / scale)))
        wt = max(wt, float(np.max(np.abs(t - rt))))
    print(f"{os.path.basename(path)}: {2 * nd} matrices, max |factors - oracle| / max|factors| = {worst:.3e}, max |tau - oracle| = {wt:.3e}",
          "OK" if worst < 1e-12 and wt < 1e-12 else "MISMATCH")
