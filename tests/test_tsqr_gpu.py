"""K5 parity: TSQR R factor vs the oracle's qrBlocked! R after row-phase normalisation
(DESIGN.md "TSQR sign": the tree sees different pivots than sequential Householder, so row signs may
differ; |R| and R up to D = diag(+-1) must agree)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _normalise(R):
    d = np.sign(np.diag(R)).copy()
    d[d == 0] = 1.0
    return R * (-d)[:, None]          # reference convention: negative diagonal for positive pivots


@pytest.mark.parametrize("m,n", [(4096, 64), (100000, 64), (5000, 32), (777, 17), (64, 64), (65, 64), (10, 4), (3, 5),
                                 (300000, 8), (1000003, 64), (40001, 33), (257, 64), (256, 40), (37888, 64)])
def test_tsqr_matches_oracle_up_to_row_signs(gla, oracle, m, n):
    rng = np.random.default_rng(m + n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    R = gla.tsqr_R(A)
    assert np.array_equal(np.tril(R, -1), np.zeros((n, n)))
    ref_f, _ = oracle.qr_blocked(A, 12)
    k = min(m, n)
    Rref = np.zeros((n, n))
    Rref[:k] = np.triu(ref_f)[:k]
    a, b = _normalise(R), _normalise(Rref)
    if m >= n:
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
    # Gram identity holds regardless of signs
    G = A.T @ A
    assert np.max(np.abs(R.T @ R - G)) <= 1e-12 * np.max(np.abs(G)) * max(1, m // 1000)


def test_tsqr_host_path_streams_row_chunks(gla, oracle):
    """m > 1.5 * 2^20 rows: the host-pointer entry point pipelines 2^20-row chunks (upload / reduce / stack of R factors /
    final fold) instead of uploading the whole matrix; 3,500,003 x 16 has a ragged last chunk and uses every ring slot."""
    m, n = 3500003, 16
    rng = np.random.default_rng(17)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    R = gla.tsqr_R(A)
    assert np.array_equal(np.tril(R, -1), np.zeros((n, n)))
    ref_f, _ = oracle.qr_blocked(A, 12)
    a, b = _normalise(R), _normalise(np.triu(ref_f)[:n])
    assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))


def test_tsqr_sharded_combine(gla, oracle):
    """The multi-GPU path on one device: 4 row shards -> local R -> stacked combine == single-shot R."""
    import torch
    m, n, G = 1 << 16, 64, 4
    g = torch.Generator(device="cuda").manual_seed(123)
    A = torch.randn((n, m), generator=g, device="cuda", dtype=torch.float64)  # column-major m x n
    Rs = torch.zeros((G, n, n), device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    rows = m // G
    for r in range(G):
        gla.tsqr_local_dev(A.data_ptr() + r * rows * 8, rows, n, m, Rs[r].data_ptr(), n, st)
    R = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    gla.tsqr_combine_dev(Rs.data_ptr(), G, n, R.data_ptr(), n, st)
    torch.cuda.synchronize()
    Rh = R.cpu().numpy().T   # stored column-major
    Ah = np.asfortranarray(A.cpu().numpy().T)
    ref_f, _ = oracle.qr_blocked(Ah, 12)
    a, b = _normalise(Rh), _normalise(np.triu(ref_f)[:n])
    assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))


@pytest.mark.parametrize("count,n", [(1, 64), (3, 17), (8, 64), (40, 33)])
def test_tsqr_combine_of_stacked_blocks(gla, oracle, count, n):
    """gla_dtsqr_combine_dev on `count` stacked n x n upper factors == QR of the stacked matrix."""
    import torch
    rng = np.random.default_rng(count * 100 + n)
    # well-conditioned upper factors (R of random tall blocks); a random triangular matrix would have kappa ~ 2^n
    Rs = np.stack([np.linalg.qr(rng.standard_normal((4 * n, n)))[1] for _ in range(count)])   # block b, row i, col j
    dRs = torch.from_numpy(np.ascontiguousarray(np.transpose(Rs, (0, 2, 1)))).cuda()   # each block column-major, ld n
    R = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    gla.tsqr_combine_dev(dRs.data_ptr(), count, n, R.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    Rh = R.cpu().numpy().T
    ref_f, _ = oracle.qr_blocked(np.asfortranarray(Rs.reshape(count * n, n)), 12)
    a, b = _normalise(Rh), _normalise(np.triu(ref_f)[:n])
    assert np.array_equal(np.tril(Rh, -1), np.zeros((n, n)))
    assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))


def test_tsqr_allreduce_entry_point_single_rank(gla, oracle):
    """gla_nccl_* + gla_dtsqr_allreduce_dev with a one-rank communicator (the N>1 exchange path of bench.py, minus
    the peers): local R -> ncclAllGather -> stack reduction must reproduce the single-shot R."""
    import torch
    m, n = 50000, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((n, m), generator=g, device="cuda", dtype=torch.float64)   # column-major m x n
    st = torch.cuda.current_stream().cuda_stream
    Rloc = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    stack = torch.zeros((1, n, n), device="cuda", dtype=torch.float64)
    R = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    gla.tsqr_local_dev(A.data_ptr(), m, n, m, Rloc.data_ptr(), n, st)
    comm = gla.TsqrComm(0, 1, lambda raw: raw)
    try:
        comm.allreduce_R(Rloc.data_ptr(), n, stack.data_ptr(), R.data_ptr(), n, st)
        torch.cuda.synchronize()
    finally:
        comm.destroy()
    ref_f, _ = oracle.qr_blocked(np.asfortranarray(A.cpu().numpy().T), 12)
    a, b = _normalise(R.cpu().numpy().T), _normalise(np.triu(ref_f)[:n])
    assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
