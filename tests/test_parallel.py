"""
CPU tests of the multi-GPU host logic under gloo, world_size 2: row partitioning,
the padded all-gather of factor slices, and the cost-dealt kNN row partition.
The per-shard compute is done by the oracle here (no GPU), which checks that
sharding + exchange reproduces the single-process result exactly.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from lkpy_b200 import data
from lkpy_b200.parallel import allgather_rows, deal_by_cost, row_bounds_by_nnz, shard_csr

from helpers import small_synth


def test_row_bounds_balanced():
    inter = small_synth(2000, 700, 60000, seed=2)
    ui, iu = data.als_implicit_matrices(inter)
    for world in (1, 2, 3, 8):
        for csr in (ui, iu):
            b = row_bounds_by_nnz(csr.indptr, world)
            assert b[0] == 0 and b[-1] == csr.shape[0] and np.all(np.diff(b) >= 0)
            per = np.diff(csr.indptr[b].astype(np.int64))
            assert per.sum() == csr.nnz
            biggest_row = np.diff(csr.indptr).max()
            assert per.max() - per.min() <= 2 * biggest_row
            sh = shard_csr(csr, int(b[0]), int(b[1]))
            assert sh.nnz == per[0] and sh.shape[1] == csr.shape[1]


def test_deal_by_cost():
    rng = np.random.default_rng(0)
    cost = (rng.pareto(1.2, size=5000) * 100).astype(np.int64)
    for world in (2, 4, 8):
        parts = deal_by_cost(cost, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(5000))
        tot = np.array([cost[p].sum() for p in parts], dtype=np.float64)
        assert tot.max() / tot.mean() < 1.0 + cost.max() / tot.mean() + 0.05


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inter = small_synth(700, 300, 15000, seed=4)
        ui, iu = data.als_implicit_matrices(inter)
        rng = np.random.default_rng(1)
        k = 16
        p = (rng.standard_normal((inter.n_users, k)) * 0.1).astype(np.float32)
        q = (rng.standard_normal((inter.n_items, k)) * 0.1).astype(np.float32)
        P, Q = torch.from_numpy(p.copy()), torch.from_numpy(q.copy())
        for csr, this, other in ((ui, P, Q), (iu, Q, P)):
            o32, _ = oracle.otor(other.numpy(), 0.1)
            ref, dref = oracle.als_half("implicit", csr, this.numpy(), other.numpy(), otor_mat=o32)
            b = row_bounds_by_nnz(csr.indptr, world)
            lo, hi = int(b[rank]), int(b[rank + 1])
            new, d = oracle.als_half(
                "implicit", shard_csr(csr, lo, hi), this.numpy()[lo:hi], other.numpy(), otor_mat=o32
            )
            this[lo:hi] = torch.from_numpy(new)
            allgather_rows(this, b, rank, world)
            assert np.array_equal(this.numpy(), ref)
            d2 = torch.tensor([d * d], dtype=torch.float64)
            dist.all_reduce(d2)
            assert float(d2.sqrt()) == pytest.approx(dref, rel=1e-9)
        # kNN: rows dealt by cost, disjoint fixed-width rows summed = gathered
        kui, kiu, _ = data.knn_item_matrices(inter, True)
        nu = np.diff(kui.indptr)
        cost = np.array([nu[kiu.indices[kiu.indptr[i] : kiu.indptr[i + 1]]].sum() for i in range(inter.n_items)])
        mine = deal_by_cost(cost, world)[rank]
        K = 5
        full = oracle.knn_build(kui, kiu, 1e-6, K)
        cols = torch.zeros((inter.n_items, K), dtype=torch.int32)
        vals = torch.zeros((inter.n_items, K), dtype=torch.float32)
        cnt = torch.zeros(inter.n_items, dtype=torch.int32)
        for i in mine:
            row = oracle.knn_build(kui, kiu, 1e-6, K, rows=(int(i), int(i) + 1))
            n = row.nnz
            cols[i, :n] = torch.from_numpy(row.indices)
            vals[i, :n] = torch.from_numpy(row.data)
            cnt[i] = n
        for t in (cols, vals, cnt):
            dist.all_reduce(t)
        assert np.array_equal(cnt.numpy(), np.diff(full.indptr))
        m = np.arange(K)[None, :] < cnt.numpy()[:, None]
        assert np.array_equal(cols.numpy()[m], full.indices)
        assert np.array_equal(vals.numpy()[m].view(np.int32), full.data.view(np.int32))
    finally:
        dist.destroy_process_group()


def test_sharded_paths_world2():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def _worker_checks(rank: int, world: int, port: int):
    """The cross-rank checks of the scale-out bench (bench_scale.py) on CPU tensors under gloo."""
    import sys
    from pathlib import Path

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
        import bench_scale

        dev = torch.device("cpu")
        g = torch.Generator().manual_seed(5)
        u = torch.randint(0, 1000, (5000,), generator=g, dtype=torch.int32)
        i = torch.randint(0, 300, (5000,), generator=g, dtype=torch.int32)
        r = torch.rand(5000, generator=g)
        # every rank holds the same matrix / the same result: both checks pass
        assert bench_scale._inputs_equal_across_ranks([u, i, r], dev, world)
        assert bench_scale._equal_across_ranks([u.float(), r], dev, world)
        # one rating differs in its last bit on one rank (what a non-reproducible device prefix sum did to the
        # generator): both checks must see it
        r2 = r.clone()
        if rank == 1:
            r2.view(torch.int32)[1234] ^= 1
        assert not bench_scale._inputs_equal_across_ranks([u, i, r2], dev, world)
        assert not bench_scale._equal_across_ranks([r2], dev, world)
        # user-sharded scoring: contiguous user ranges balanced by history length cover every user once
        hist = np.concatenate([[0], np.cumsum(np.random.default_rng(3).integers(0, 400, 3000))])
        b = row_bounds_by_nnz(hist, world)
        mine = torch.zeros(3000, dtype=torch.int32)
        mine[int(b[rank]) : int(b[rank + 1])] = 1
        dist.all_reduce(mine)
        assert bool((mine == 1).all())
    finally:
        dist.destroy_process_group()


def test_cross_rank_checks_world2():
    mp.spawn(_worker_checks, args=(2, _free_port()), nprocs=2, join=True)
