"""Compare the dumps written by tools/sweep_batched against the CPU oracle (test infrastructure).
usage: python tools/sweep_check.py gpurun_out/sweep_v*.bin

Dump layout (Float64): two parts (first / last NDUMP matrices of the batch), each = inputs (NDUMP x 32 x 32, column-major
per matrix), factors (same shape) and tau (NDUMP x 32)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TOL = 1e-12


def check(path):
    """-> (number of matrices, max |factors - oracle| / max|factors| per matrix, max |tau - oracle|, ok)."""
    from oracle import oracle

    raw = np.fromfile(path, dtype=np.float64)
    nd = raw.size // (2 * (2 * 1024 + 32))
    if nd == 0 or raw.size != nd * 2 * (2 * 1024 + 32):
        raise ValueError(f"{path}: not a sweep_batched dump ({raw.size} doubles)")
    worst = 0.0
    wt = 0.0
    off = 0
    for _part in range(2):
        a = raw[off:off + nd * 1024].reshape(nd, 32, 32); off += nd * 1024      # (batch, col, row)
        f = raw[off:off + nd * 1024].reshape(nd, 32, 32); off += nd * 1024
        t = raw[off:off + nd * 32].reshape(nd, 32); off += nd * 32
        rf, rt = oracle.qr_batched(np.transpose(a, (0, 2, 1)))
        rf = np.transpose(rf, (0, 2, 1))
        scale = np.max(np.abs(rf), axis=(1, 2), keepdims=True)
        worst = max(worst, float(np.max(np.abs(f - rf) / scale)))
        wt = max(wt, float(np.max(np.abs(t - rt))))
    return 2 * nd, worst, wt, bool(worst < TOL and wt < TOL)


def main(paths):
    bad = 0
    for path in paths:
        n, worst, wt, ok = check(path)
        print(f"{os.path.basename(path)}: {n} matrices, max |factors - oracle| / max|factors| = {worst:.3e}, "
              f"max |tau - oracle| = {wt:.3e}", "OK" if ok else "MISMATCH")
        bad += not ok
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
