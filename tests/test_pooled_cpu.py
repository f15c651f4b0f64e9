"""CPU, world_size 2 over gloo: the host-side logic of the pooled (sharded) mode.

The kernels need a GPU, so what runs here is the partition itself with the oracle standing in for
the kernels: every rank computes the row statistics / gradient rows of ITS row range against all keys,
the ranks exchange them (here with gloo collectives; the product moves the same row ranges with its own kernels over
NVLink peer memory, csrc/xchg.cu), and the result must equal the single-process oracle."""
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
    # gradient rows: every rank holds the complete rows of its 128-aligned block; in the product the scatter of a rank
    # PULLS row i of a local pixel from rank i // per (mscs_scatter_sectors_pull).  Emulated with an all-gather of
    # the blocks: the rule "row i lives on rank i // per" must reproduce every row, for every owner of pixels
    from mscs_b200 import TorchDistComm
    comm = TorchDistComm()
    assert comm.world == world and comm.rank == rank
    Cp = 8
    per = ((N + 127) // 128 + world - 1) // world * 128
    rows_full = torch.from_numpy(rng.randn(world * per, Cp)).float()          # same on every rank (same seed)
    mine_blk = rows_full[rank * per:(rank + 1) * per].contiguous()
    blocks = comm.all_gather(mine_blk).view(world, per, Cp)
    rows_i = torch.arange(N)
    pulled = blocks[rows_i // per, rows_i % per]
    ok_blocks = torch.equal(pulled, rows_full[:N]) and shard_rows(N, world, rank) == (min(N, rank * per), min(N, (rank + 1) * per))
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


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_backward_needs_only_row_statistics(world):
    """The algebra behind the pooled backward, in numpy with the oracle as
    the single-process answer: once the row statistics (neg, S, P) of ALL rows are known on every rank, a rank forms
    both G_ij (its rows as anchors) and G_ji (its rows as keys) itself, so the gradient rows of its 128-aligned row
    range need no reduce-scatter -- single-scale: dF_i = sum_j (G_ij + G_ji) f_j / tau; cross-scale: dA for its anchor
    rows, dK for its key rows.  Concatenating the ranks' row blocks must give the oracle's full gradients."""
    from mscs_b200 import shard_rows
    from oracle import loss_fp64
    rng = np.random.RandomState(world)

    def unit(n, c):
        f = rng.randn(n, c)
        return f / np.linalg.norm(f, axis=1, keepdims=True)

    def G_block(Fa, ya, Fk, yk, st, tau, self_mask, rows_a, N1):
        """G[rows_a, :] from the row statistics of those anchor rows."""
        E = np.exp(Fa[rows_a] @ Fk.T / tau)
        pos = ya[rows_a, None] == yk[None, :]
        negm = ~pos
        if self_mask:
            pos[np.arange(len(rows_a)), rows_a] = False
        P = st["P"][rows_a]
        div = P if self_mask else np.where(P > 0, P, 1.0)
        inv = (1.0 / (div * N1))[:, None]
        n_i = st["neg"][rows_a, None]
        return np.where(pos, -inv * n_i / (E + n_i), 0.0) + np.where(negm, inv * st["S"][rows_a, None] * E, 0.0)

    # single-scale term: 1000 rows, 7 classes, class-sorted like the kernel layout
    N, C, tau = 1000, 24, 0.1
    F, y = unit(N, C), np.sort(rng.randint(0, 7, N))
    _, da, dk, st = loss_fp64.term(F, y, F, y, tau, True)
    want = da + dk
    got = np.zeros_like(want)
    allrows = np.arange(N)
    for r in range(world):
        b, e = shard_rows(N, world, r)
        rows = np.arange(b, e)
        if len(rows) == 0:
            continue
        G_rows = G_block(F, y, F, y, st, tau, True, rows, N)                 # G_ij, i in my rows
        G_cols = G_block(F, y, F, y, st, tau, True, allrows, N)[:, rows]     # G_ji, i in my rows (from j's statistics)
        got[b:e] = (G_rows + G_cols.T) @ F / tau
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())

    # cross-scale term: 900 anchors x 300 keys, some anchors without any positive (class 6 absent from the keys)
    N1, N2, tau = 900, 300, 0.07
    Fa, ya = unit(N1, C), np.sort(rng.randint(0, 7, N1))
    Fk, yk = unit(N2, C), np.sort(rng.randint(0, 6, N2))
    _, da, dk, st = loss_fp64.term(Fa, ya, Fk, yk, tau, False)
    assert (st["P"] == 0).any()
    got_a, got_k = np.zeros_like(da), np.zeros_like(dk)
    G_full = G_block(Fa, ya, Fk, yk, st, tau, False, np.arange(N1), N1)
    for r in range(world):
        b, e = shard_rows(N1, world, r)
        if e > b:
            got_a[b:e] = G_block(Fa, ya, Fk, yk, st, tau, False, np.arange(b, e), N1) @ Fk / tau
        kb, ke = shard_rows(N2, world, r)
        if ke > kb:
            got_k[kb:ke] = G_full[:, kb:ke].T @ Fa / tau                     # columns of G from the exchanged statistics
    assert np.abs(got_a - da).max() <= 1e-12 * max(1.0, np.abs(da).max())
    assert np.abs(got_k - dk).max() <= 1e-12 * max(1.0, np.abs(dk).max())
