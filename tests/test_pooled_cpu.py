"""CPU, world_size 2 over gloo: the host-side logic of the pooled (sharded) mode.

The kernels need a GPU, so what runs here is the partition itself with the oracle standing in for
the kernels: every rank computes the row statistics / gradient rows of ITS row range against all keys,
the ranks exchange them with the same collectives the product uses (sum of disjoint supports), and
the result must equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mscs_b200 import shard_rows
    from oracle import loss_fp64
    rng = np.random.RandomState(0)
    N, Cc = 700, 16
    F = rng.randn(N, Cc)
    F /= np.linalg.norm(F, axis=1, keepdims=True)
    y = np.sort(rng.randint(0, 6, N))
    b, e = shard_rows(N, world, rank)
    assert b % 128 == 0 and (e % 128 == 0 or e == N)
    # forward: statistics of my rows only, zeros elsewhere, summed over ranks
    _, _, _, st = loss_fp64.term(F[b:e], y[b:e], F, y, 0.1, False, need_grad=False)
    # (self-mask handled by index offset: recompute with the diagonal of the global matrix)
    L = F[b:e] @ F.T / 0.1
    E = np.exp(L)
    neg = (E * (y[b:e, None] != y[None, :])).sum(1)
    stats = torch.zeros(N, dtype=torch.float64)
    stats[b:e] = torch.from_numpy(neg)
    dist.all_reduce(stats)
    full = np.exp(F @ F.T / 0.1)
    want = (full * (y[:, None] != y[None, :])).sum(1)
    ok_stats = np.allclose(stats.numpy(), want, rtol=1e-12)
    # ranges tile [0, N) exactly once
    cover = torch.zeros(N, dtype=torch.int64)
    cover[b:e] += 1
    dist.all_reduce(cover)
    ok_cover = bool((cover == 1).all())
    # counts all-gather: image-major concatenation of contiguous per-rank blocks
    mine = torch.full((3, 5), rank, dtype=torch.int32)
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    ok_gather = all(int(allc[r][0, 0]) == r for r in range(world))
    # gradient rows: every rank holds the complete rows of its 128-aligned block; the product exchanges them with an
    # in-place block all-gather (TorchDistComm.all_gather_blocks_async) instead of an all-reduce of a zero-filled buffer
    from mscs_b200 import TorchDistComm
    comm = TorchDistComm()
    Cp = 8
    per = ((N + 127) // 128 + world - 1) // world * 128
    rows_full = torch.from_numpy(rng.randn(world * per, Cp)).float()          # same on every rank (same seed)
    buf = torch.zeros(world * per * Cp + 3 * Cp)                               # slack like the dF slab
    buf[rank * per * Cp:(rank + 1) * per * Cp] = rows_full[rank * per:(rank + 1) * per].reshape(-1)
    comm.all_gather_blocks_async(buf[:world * per * Cp], per * Cp).wait()
    ok_blocks = torch.equal(buf[:world * per * Cp], rows_full.reshape(-1)) and float(buf[world * per * Cp:].abs().max()) == 0.0
    out[rank] = (ok_stats, ok_cover, ok_gather and ok_blocks)
    dist.destroy_process_group()


def test_shard_rows_properties():
    from mscs_b200 import shard_rows
    for N in (1, 127, 128, 129, 700, 9804, 64448):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                b, e = shard_rows(N, world, r)
                assert b == seen and b <= e <= N
                assert b % 128 == 0 or b == N
                seen = e
            assert seen == N


@pytest.mark.timeout(120)
def test_pooled_partition_gloo_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        assert out[r] == (True, True, True), (r, out[r])
