"""Small problems through every kernel added in round 2 (for compute-sanitizer memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
g = ge.load()
rng = np.random.default_rng(1)
# Cholesky panel path (two outer blocks, ragged), LDLt both triangles
for n in (200, 449):
    X = rng.standard_normal((n, n)); S = np.asfortranarray(X.T @ X + n * np.eye(n))
    L = np.tril(g.cholRecursive_(S.copy(order="F")))
    assert np.abs(L @ L.T - S).max() < 1e-9 * n
    for uplo in ("L", "U"):
        g.ldlt_(S.copy(order="F"), uplo)
# cluster panel kernel + doubling larft + wide ormqr / thin Q, f64 and f32 (3xTF32 contraction)
for dt in (np.float64, np.float32):
    A = np.asfortranarray(rng.standard_normal((700, 500)).astype(dt))
    qr = g.qrBlocked_(A.copy(order="F"))
    Q = qr.thinQ()
    assert np.abs(Q.T @ Q - np.eye(500)).max() < (1e-3 if dt == np.float32 else 1e-10)
print("san ok")
