"""Reproducibility under concurrency.  Every reduction in the library runs in a fixed order, so repeated runs on the
same input must be BITWISE equal -- also while the look-ahead stream of the blocked QR is active and while unrelated
kernels run on other streams.  (Regression test for the generic-load-of-TMA-written-shared-memory bug, DESIGN.md.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n", [("d", 2304), ("z", 1600), ("d", 4096)])
def test_blocked_qr_bitwise_reproducible_with_lookahead(gla, kind, n):
    import torch
    dt = torch.complex128 if kind == "z" else torch.float64
    npdt = np.complex128 if kind == "z" else np.float64
    g = torch.Generator(device="cuda").manual_seed(n)
    src = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64).to(dt)
    if kind == "z":
        src = src + 1j * torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
    tau = torch.zeros(n, device="cuda", dtype=dt)
    st = torch.cuda.current_stream().cuda_stream
    ref = None
    for _ in range(6):
        dA = src.clone()
        gla.qr_blocked_dev(dA.data_ptr(), n, n, n, tau.data_ptr(), 0, st, npdt)   # n > 512: look-ahead active
        torch.cuda.synchronize()
        if ref is None:
            ref = dA
            A0, R = src.t(), torch.triu(dA.t())          # storage is column-major: dA[j, i] = F[i, j]
            G = A0.conj().t() @ A0
            assert ((R.conj().t() @ R - G).abs().max() / G.abs().max()).item() < 1e-12
        else:
            assert torch.equal(torch.view_as_real(dA) if kind == "z" else dA,
                               torch.view_as_real(ref) if kind == "z" else ref)


def test_cholesky_bitwise_reproducible_under_foreign_streams(gla):
    """The TMA/DMMA contraction chain while a high-priority stream floods the GPU with unrelated kernels."""
    import torch
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
    S = X.t() @ X + n * torch.eye(n, device="cuda", dtype=torch.float64)
    info = torch.zeros(1, device="cuda", dtype=torch.int32)
    main = torch.cuda.Stream()
    noise = torch.cuda.Stream(priority=-1)
    nz = [torch.zeros(1 << 22, device="cuda") for _ in range(4)]
    big = torch.zeros((2048, 2048), device="cuda")
    ref = None
    torch.cuda.synchronize()
    for _ in range(8):
        dA = S.clone()
        torch.cuda.synchronize()
        with torch.cuda.stream(main):
            gla.chol_recursive_dev(dA.data_ptr(), n, n, info.data_ptr(), 1, main.cuda_stream)
        with torch.cuda.stream(noise):
            for k in range(200):
                nz[k & 3].add_(1.0)
                if k % 40 == 0:
                    torch.mm(big, big)
        torch.cuda.synchronize()
        assert int(info.item()) == 0
        if ref is None:
            ref = dA
            L = torch.tril(dA.t())
            assert ((L @ L.t() - S).norm() / S.norm()).item() < 1e-13
        else:
            assert torch.equal(torch.tril(dA.t()), torch.tril(ref.t()))


def _noise(torch, stream, nz, big, rounds=150):
    with torch.cuda.stream(stream):
        for k in range(rounds):
            nz[k & 3].add_(1.0)
            if k % 40 == 0:
                torch.mm(big, big)


def test_batched_and_tsqr_bitwise_reproducible_under_foreign_streams(gla):
    import torch
    g = torch.Generator(device="cuda").manual_seed(11)
    noise = torch.cuda.Stream(priority=-1)
    main = torch.cuda.Stream()
    nz = [torch.zeros(1 << 22, device="cuda") for _ in range(4)]
    big = torch.zeros((2048, 2048), device="cuda")
    # batched 32x32
    batch = 1 << 16
    src = torch.randn((batch, 32, 32), generator=g, device="cuda", dtype=torch.float64)
    tau = torch.zeros((batch, 32), device="cuda", dtype=torch.float64)
    ref = ref_tau = None
    for _ in range(4):
        dA = src.clone()
        torch.cuda.synchronize()
        with torch.cuda.stream(main):
            gla.qr_batched_dev(dA.data_ptr(), 32, 32, batch, tau.data_ptr(), main.cuda_stream)
        _noise(torch, noise, nz, big)
        torch.cuda.synchronize()
        if ref is None:
            ref, ref_tau = dA, tau.clone()
        else:
            assert torch.equal(dA, ref) and torch.equal(tau, ref_tau)
    # TSQR
    m, n = 1 << 20, 64
    A = torch.randn((n, m), generator=g, device="cuda", dtype=torch.float64)
    R = torch.zeros((n, n), device="cuda", dtype=torch.float64)
    refR = None
    for _ in range(4):
        R.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(main):
            gla.tsqr_local_dev(A.data_ptr(), m, n, m, R.data_ptr(), n, main.cuda_stream)
        _noise(torch, noise, nz, big)
        torch.cuda.synchronize()
        if refR is None:
            refR = R.clone()
            Ru = torch.triu(R.t())
            G = A @ A.t()
            assert ((Ru.t() @ Ru - G).abs().amax() / G.abs().amax()).item() < 1e-12
        else:
            assert torch.equal(R, refR)


def test_float32_qr_on_tcgen05_bitwise_reproducible_under_foreign_streams(gla):
    """Float32 qrBlocked! (contractions on the tcgen05 kernel: TMA-written tiles rewritten by the split warps through the
    generic proxy and read by the tensor core through the async proxy) with look-ahead active and a high-priority stream
    flooding the GPU: bitwise equal run to run, Gram identity at Float32 accuracy."""
    import torch
    n = 2304
    g = torch.Generator(device="cuda").manual_seed(23)
    src = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float32)
    tau = torch.zeros(n, device="cuda", dtype=torch.float32)
    main = torch.cuda.Stream()
    noise = torch.cuda.Stream(priority=-1)
    nz = [torch.zeros(1 << 22, device="cuda") for _ in range(4)]
    big = torch.zeros((2048, 2048), device="cuda")
    ref = None
    torch.cuda.synchronize()
    for it in range(6):
        dA = src.clone()
        torch.cuda.synchronize()
        with torch.cuda.stream(main):
            gla.qr_blocked_dev(dA.data_ptr(), n, n, n, tau.data_ptr(), 0, main.cuda_stream, np.float32)
        if it > 0:
            _noise(torch, noise, nz, big)
        torch.cuda.synchronize()
        if ref is None:
            ref = dA
            A0, R = src.t().double(), torch.triu(dA.t()).double()
            G = A0.t() @ A0
            assert ((R.t() @ R - G).abs().max() / G.abs().max()).item() < 2e-5
        else:
            assert torch.equal(dA, ref)


def test_twosided_bitwise_reproducible(gla):
    """The persistent two-sided kernels: redundant per-CTA reflectors and fixed-order sums -> bitwise equal run to run,
    also with an unrelated stream competing for the SMs while the cooperative grid is resident."""
    import torch
    n = 700
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
    t1 = torch.zeros(n, device="cuda", dtype=torch.float64)
    t2 = torch.zeros(n, device="cuda", dtype=torch.float64)
    main = torch.cuda.Stream()
    noise = torch.cuda.Stream()
    nz = [torch.zeros(1 << 20, device="cuda") for _ in range(4)]
    big = torch.zeros((1024, 1024), device="cuda")
    ref = None
    torch.cuda.synchronize()
    for it in range(4):
        dA = src.clone()
        torch.cuda.synchronize()
        with torch.cuda.stream(main):
            gla.bidiagonalize_dev(dA.data_ptr(), n, n, n, t1.data_ptr(), t2.data_ptr(), main.cuda_stream)
        if it > 1:
            _noise(torch, noise, nz, big, rounds=40)
        torch.cuda.synchronize()
        if ref is None:
            ref = (dA, t1.clone(), t2.clone())
        else:
            assert torch.equal(dA, ref[0]) and torch.equal(t1, ref[1]) and torch.equal(t2, ref[2])


@pytest.mark.parametrize("m,n", [(2304, 2304), (3000, 1700), (1000, 2600)])
def test_host_pointer_qr_with_streamed_upload_equals_device_resident(gla, m, n):
    """gla_dgeqr_blocked on host memory: from n > 1536 (and m > 768) the matrix arrives in column chunks while the first outer
    block is factorised and the first far update is cut along the chunk boundaries -- with the split-K slicing of the whole,
    so the result must be BITWISE the device-resident one.  (1000 x 2600: wide, a single outer block row range.)"""
    import torch
    g = torch.Generator(device="cuda").manual_seed(m + n)
    src = torch.randn((n, m), generator=g, device="cuda", dtype=torch.float64)   # storage of a column-major m x n matrix
    k = min(m, n)
    dA = src.clone()
    dtau = torch.zeros(k, device="cuda", dtype=torch.float64)
    gla.qr_blocked_dev(dA.data_ptr(), m, n, m, dtau.data_ptr(), 0, torch.cuda.current_stream().cuda_stream, np.float64)
    torch.cuda.synchronize()
    for pinned in (True, False):
        hA = torch.empty((n, m), dtype=torch.float64, pin_memory=pinned)
        htau = torch.zeros(k, dtype=torch.float64, pin_memory=pinned)
        hA.copy_(src)
        torch.cuda.synchronize()
        gla.qr_blocked_ptr(hA.data_ptr(), m, n, m, htau.data_ptr(), 0, np.float64)
        assert torch.equal(hA.cuda(), dA)
        assert torch.equal(htau.cuda(), dtau)
