"""Small problems through the kernels added late in round 2 (tcgen05 Float32 contraction, two-sided reductions,
per-warp TSQR, the st.async exchange of the cluster panel kernel, the warp-per-matrix batched kernel) for
compute-sanitizer memcheck / racecheck / synccheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
g = ge.load()
rng = np.random.default_rng(2)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "umma"):
    # tcgen05 kernel directly: rank-k update with a full tile + ragged edges, and through a Float32 factorisation
    for (n, k) in ((128, 40), (300, 200)):
        A = np.asfortranarray(rng.standard_normal((n, k)).astype(np.float32))
        C0 = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32))
        C = C0.copy(order="F")
        g.rankUpdate_(C, A, -1.0)
        ref = C0.astype(np.float64) - A.astype(np.float64) @ A.astype(np.float64).T
        assert np.abs(np.tril(C) - np.tril(ref)).max() < 1e-4
    A = np.asfortranarray(rng.standard_normal((900, 900)).astype(np.float32))
    qr = g.qrBlocked_(A.copy(order="F"))
    R = np.triu(qr.factors).astype(np.float64)
    G = A.astype(np.float64).T @ A.astype(np.float64)
    assert np.abs(R.T @ R - G).max() / np.abs(G).max() < 1e-4
if which in ("all", "twosided"):
    for dt in (np.float64, np.complex128, np.float32):
        A = rng.standard_normal((70, 50)).astype(dt)
        if dt == np.complex128:
            A = A + 1j * rng.standard_normal((70, 50))
        g.bidiagonalize_(np.asfortranarray(A))
        g.bidiagonalize_(np.asfortranarray(A.T.copy()))
        B = np.asfortranarray(A[:50, :50])
        g.hessenberg_(B.copy(order="F"))
        S = np.asfortranarray(B + B.conj().T)
        g.symtri_(S.copy(order="F"), "L")
        g.symtri_(S.copy(order="F"), "U")
if which in ("all", "tsqr"):
    m = 148 * 8 * 256 + 77
    A = np.asfortranarray(rng.standard_normal((m, 24)))
    R = g.tsqr_R(A)
    G = A.T @ A
    assert np.abs(R.T @ R - G).max() / np.abs(G).max() < 1e-12
if which in ("all", "cluster"):
    # cluster panel kernel (st.async exchange on transaction barriers): 1 ... 8 CTAs per cluster, ragged last slab
    for dt, shapes in ((np.float64, ((70, 70), (300, 130), (1000, 200))), (np.complex128, ((260, 100),)), (np.float32, ((520, 70),))):
        for (m, n) in shapes:
            A = rng.standard_normal((m, n)).astype(dt)
            if dt == np.complex128:
                A = A + 1j * rng.standard_normal((m, n))
            qr = g.qrBlocked_(np.asfortranarray(A))
            R = np.triu(qr.factors[:n, :]).astype(np.complex128 if dt == np.complex128 else np.float64)
            A64 = A.astype(R.dtype)
            G = A64.conj().T @ A64
            assert np.abs(R.conj().T @ R - G).max() / np.abs(G).max() < (1e-4 if dt == np.float32 else 1e-12)
if which in ("all", "batched"):
    for dt in (np.complex128, np.float64):
        A = rng.standard_normal((37, 24, 20)).astype(dt)   # 37 matrices 24 x 20, each stored column-major
        g.qr_batched_(np.array(np.transpose(A, (0, 2, 1)), order="C", copy=True))
print("san ok")
