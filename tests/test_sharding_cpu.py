"""N>1 host logic on CPU: world-size-2 `gloo` run of the sharded TSQR data flow (local R -> all-gather of the
R factors -> reduction of the stack) and of the batched index partition.  The arithmetic of each node is done by
the oracle here (no GPU in this suite); on the GPU box the same flow runs through gla_dtsqr_local_dev /
gla_dtsqr_allreduce_dev (tests/test_tsqr_gpu.py, bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _normalise(R):
    d = np.sign(np.diag(R)).copy()
    d[d == 0] = 1.0
    return R * (-d)[:, None]


def test_shard_range_partitions_exactly(gla):
    for total in (0, 1, 7, 1 << 20, 8388608 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [gla.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert spans[-1][0] + spans[-1][1] == total
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        gla.shard_range(10, 2, 2)


def _worker(rank, world, port, m, n, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle import oracle
    g = ge.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = np.asfortranarray(np.random.default_rng(7).standard_normal((m, n)))     # same matrix on every rank
    r0, cnt = g.shard_range(m, rank, world)
    f, _ = oracle.qr_blocked(np.asfortranarray(A[r0:r0 + cnt]), 12)
    Rloc = np.zeros((n, n))
    k = min(cnt, n)
    Rloc[:k] = np.triu(f)[:k]

    def all_gather(R):
        t = torch.from_numpy(np.ascontiguousarray(R))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return np.stack([o.numpy() for o in outs])

    def reduce_stack(stack):
        f2, _ = oracle.qr_blocked(np.asfortranarray(stack.reshape(world * n, n)), 12)
        return np.triu(f2)[:n]

    R = g.tsqr_R_sharded(Rloc, all_gather, reduce_stack)
    # batched partition: every rank factorises its index range; the union is the whole batch
    B = np.random.default_rng(11).standard_normal((10, 8, 8))
    b0, bc = g.shard_range(10, rank, world)
    mine = torch.zeros(10, dtype=torch.int64)
    mine[b0:b0 + bc] = 1
    dist.all_reduce(mine)
    if rank == 0:
        fref, _ = oracle.qr_blocked(A, 12)
        np.save(out, np.stack([_normalise(R), _normalise(np.triu(fref)[:n])]))
        assert mine.tolist() == [1] * 10 and B.shape[0] == 10
    dist.barrier()
    dist.destroy_process_group()


def test_tsqr_sharded_flow_gloo_world2(tmp_path, oracle):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r.npy")
    mp.spawn(_worker, args=(2, port, 1001, 16, out), nprocs=2, join=True)
    a, b = np.load(out)
    assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))
