"""
Multi-GPU execution of the two hot paths on one NVSwitch box: one process per
GPU, ``torch.distributed`` (NCCL) for the plumbing.

The reference has no distributed path at all; the sharding is
the natural one for an embarrassingly row-parallel map
(``implicit.rs:71-80``, ``item_train.rs:71-87``):

* **ALS** — every rank keeps full replicas of both factor tables and the row
  shard of R (user half) / Rᵀ (item half) it owns, shards being contiguous row
  ranges balanced by nonzeros.  A half-epoch solves the local rows in place in
  the replica and then all-gathers the updated slices — the single exchange step
  of the path (42 MB + 15 MB per epoch at ML-25M k=64).  ``OᵀO + λI`` is
  recomputed locally from the replica (k×k, no collective).
* **item-kNN build** — ``UI`` is replicated, item rows are dealt to ranks by
  descending cost; no collective during the build, one exchange of the
  fixed-width top-K rows at the end.

The partitioning / exchange helpers are backend-agnostic (they run under gloo
on CPU tensors in ``tests/test_parallel.py``).
"""

from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine
from .als import ALSTrainerBase, ImplicitMFTrainer
from .data import InteractionCSR


def row_bounds_by_nnz(indptr: np.ndarray, world: int) -> np.ndarray:
    """Contiguous row ranges with ~equal nonzeros: bounds[r]..bounds[r+1] belongs to rank r."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n_rows = len(indptr) - 1
    nnz = int(indptr[-1])
    targets = (np.arange(1, world, dtype=np.float64) * nnz / world).astype(np.int64)
    inner = np.searchsorted(indptr, targets, side="left")
    b = np.concatenate([[0], np.clip(inner, 0, n_rows), [n_rows]]).astype(np.int64)
    return np.maximum.accumulate(b)


def shard_csr(csr: InteractionCSR, lo: int, hi: int) -> InteractionCSR:
    a, b = int(csr.indptr[lo]), int(csr.indptr[hi])
    return InteractionCSR(
        np.ascontiguousarray(csr.indptr[lo : hi + 1] - csr.indptr[lo]).astype(csr.indptr.dtype),
        np.ascontiguousarray(csr.indices[a:b]),
        np.ascontiguousarray(csr.values[a:b]),
        (hi - lo, csr.shape[1]),
    )


def allgather_rows(full: torch.Tensor, bounds: np.ndarray, rank: int, world: int, group=None) -> None:
    """
    In-place all-gather of row slices: on entry ``full[bounds[rank]:bounds[rank+1]]``
    is current on this rank, on exit the whole of ``full`` is.  Uneven slices are
    padded to the longest one (one collective; the extra copy is D2D).
    """
    if world == 1:
        return
    sizes = np.diff(bounds)
    m = int(sizes.max())
    k = full.shape[1]
    send = torch.zeros((m, k), dtype=full.dtype, device=full.device)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    send[: hi - lo].copy_(full[lo:hi])
    recv = torch.empty((world, m, k), dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(recv.view(world * m, k), send, group=group)
    for r in range(world):
        if r == rank:
            continue
        a, b = int(bounds[r]), int(bounds[r + 1])
        full[a:b].copy_(recv[r, : b - a])


def deal_by_cost(cost: np.ndarray, world: int) -> list[np.ndarray]:
    """Deal items to ranks in descending cost, snake order: near-equal total cost per rank."""
    order = np.argsort(-np.asarray(cost), kind="stable")
    pos = np.arange(len(order))
    lap, off = pos // world, pos % world
    owner = np.where(lap % 2 == 0, off, world - 1 - off)
    return [order[owner == r] for r in range(world)]


class ShardedImplicitMFTrainer(ImplicitMFTrainer):
    """Row-sharded implicit ALS over the ranks of the default process group."""

    def __init__(self, scorer, data, options, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        super().__init__(scorer, data, options)
        k = self.config.embedding_size
        dev = self.device
        self.u_bounds = row_bounds_by_nnz(self.ui_host.indptr, self.world)
        self.i_bounds = row_bounds_by_nnz(self.iu_host.indptr, self.world)
        ulo, uhi = int(self.u_bounds[self.rank]), int(self.u_bounds[self.rank + 1])
        ilo, ihi = int(self.i_bounds[self.rank]), int(self.i_bounds[self.rank + 1])
        # replace the full matrices by this rank's row shards
        self.ui = engine.DeviceCSR.from_host(shard_csr(self.ui_host, ulo, uhi), dev)
        self.iu = engine.DeviceCSR.from_host(shard_csr(self.iu_host, ilo, ihi), dev)
        self.u_plan = engine.ALSHalfPlan.create(self.ui, k)
        self.i_plan = engine.ALSHalfPlan.create(self.iu, k)
        self.u_slice = (ulo, uhi)
        self.i_slice = (ilo, ihi)
        # every rank must start from the same factors (rank 0's draw)
        dist.broadcast(self.d_users, 0, group=group)
        dist.broadcast(self.d_items, 0, group=group)
        self.user_peers = self.item_peers = None
        if os.environ.get("LK_ALS_PEER_WRITES", "1") != "0":
            self._try_peer_tables()
        torch.cuda.empty_cache()

    def _try_peer_tables(self) -> None:
        """
        Fused exchange: put both factor tables in symmetric (peer-mapped) memory so that the solve
        kernel's epilogue stores every new row straight into all replicas over NVLink
        (``lk_als_args.d_replicas``) — the all-gather disappears into the kernel.  Falls back to
        the padded NCCL all-gather when symmetric memory is unavailable.
        """
        try:
            import torch.distributed._symmetric_memory as symm_mem

            grp = self.group if self.group is not None else dist.group.WORLD
            tabs, peers = [], []
            for t in (self.d_users, self.d_items):
                st = symm_mem.empty(tuple(t.shape), dtype=t.dtype, device=t.device)
                st.copy_(t)
                hdl = symm_mem.rendezvous(st, grp)
                off = st.data_ptr() - int(hdl.buffer_ptrs[self.rank])
                peers.append([int(hdl.buffer_ptrs[r]) + off for r in range(self.world) if r != self.rank])
                tabs.append((st, hdl))
            (self.d_users, self._hu), (self.d_items, self._hq) = tabs
            self.user_peers, self.item_peers = peers
        except Exception as e:  # noqa: BLE001
            self.user_peers = self.item_peers = None
            if self.rank == 0:
                print(f"[lkpy_b200] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather")

    def train_epoch_device(self):
        self.u_plan.status.zero_()
        self.i_plan.status.zero_()
        ulo, uhi = self.u_slice
        ilo, ihi = self.i_slice
        if self.user_peers is not None:
            # peer-write path: rows land in every replica from inside the kernel; the all-reduce of
            # the Σ‖Δ‖² scalar is the only collective and doubles as the "all shards written" barrier
            du = self._half(self.u_plan, self.d_users[ulo:uhi], self.d_items, self.d_items_bf16,
                            self.config.user_reg, replicas=self.user_peers, replica_row0=ulo)  # fmt: skip
            dist.all_reduce(du, group=self.group)
            di = self._half(self.i_plan, self.d_items[ilo:ihi], self.d_users, self.d_users_bf16,
                            self.config.item_reg, replicas=self.item_peers, replica_row0=ilo)  # fmt: skip
            dist.all_reduce(di, group=self.group)
        else:
            du = self._half(self.u_plan, self.d_users[ulo:uhi], self.d_items, self.d_items_bf16, self.config.user_reg)
            allgather_rows(self.d_users, self.u_bounds, self.rank, self.world, self.group)
            di = self._half(self.i_plan, self.d_items[ilo:ihi], self.d_users, self.d_users_bf16, self.config.item_reg)
            allgather_rows(self.d_items, self.i_bounds, self.rank, self.world, self.group)
        self.epochs_trained += 1
        return du, di

    def train_epoch_e2e(self, host_users: torch.Tensor, host_items: torch.Tensor) -> dict[str, float]:
        """
        The host-array epoch of ``ALSTrainerBase.train_epoch_e2e`` for the sharded trainer: every rank
        uploads only the rows it owns (the PCIe links work in parallel), the replicas are completed
        over NVLink, and after the epoch rank 0 reads the whole model back while the other ranks
        refresh only their own rows.  ``e2e_bytes`` holds this rank's (h2d, d2h) byte counts.
        """
        ulo, uhi = self.u_slice
        ilo, ihi = self.i_slice
        self.d_users[ulo:uhi].copy_(host_users[ulo:uhi], non_blocking=True)
        self.d_items[ilo:ihi].copy_(host_items[ilo:ihi], non_blocking=True)
        allgather_rows(self.d_users, self.u_bounds, self.rank, self.world, self.group)
        allgather_rows(self.d_items, self.i_bounds, self.rank, self.world, self.group)
        du, di = self.train_epoch_device()
        if self.user_peers is None:
            d = torch.cat([du, di])
            dist.all_reduce(d, group=self.group)
        else:
            d = torch.cat([du, di])  # already reduced (the all-reduce is the exchange barrier)
        k4 = self.d_users.shape[1] * 4
        if self.rank == 0:
            host_users.copy_(self.d_users, non_blocking=True)
            host_items.copy_(self.d_items, non_blocking=True)
            d2h = (self.d_users.shape[0] + self.d_items.shape[0]) * k4
        else:
            host_users[ulo:uhi].copy_(self.d_users[ulo:uhi], non_blocking=True)
            host_items[ilo:ihi].copy_(self.d_items[ilo:ihi], non_blocking=True)
            d2h = ((uhi - ulo) + (ihi - ilo)) * k4
        self.e2e_bytes = (((uhi - ulo) + (ihi - ilo)) * k4, d2h + 16)
        deltas = d.cpu()  # device->host read of the step's result; synchronises
        return {"deltaP": float(np.sqrt(deltas[0])), "deltaQ": float(np.sqrt(deltas[1]))}

    def train_epoch(self):
        du, di = self.train_epoch_device()
        d = torch.cat([du, di])
        if self.user_peers is None:
            dist.all_reduce(d, group=self.group)  # Σ‖Δ‖² over the shards
        st = torch.stack([self.u_plan.status, self.i_plan.status]).flatten().clone()
        dist.all_reduce(st, op=dist.ReduceOp.MAX, group=self.group)
        self._sync_host()
        if int(st.max().item()):
            raise RuntimeError("ALS solve error: a row system is not positive definite")
        self._save_user_otor()
        return {"deltaP": float(np.sqrt(d[0].item())), "deltaQ": float(np.sqrt(d[1].item()))}


def sharded_knn_build_topk(
    plan: engine.KnnBuildPlan, min_sim: float, save_nbrs: int, group=None
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Item-sharded truncated build; every rank returns the full fixed-width result."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    # deal_by_cost on the device (no host sync): plan.order is already most-expensive-first
    pos = torch.arange(plan.order.numel(), device=plan.order.device)
    lap, off = pos // world, pos % world
    owner = torch.where(lap % 2 == 0, off, world - 1 - off)
    order = plan.order[owner == rank].contiguous()
    cols, vals, cnt = plan.build_topk(min_sim, save_nbrs, order)
    # rows are disjoint across ranks: zero the padding, then a sum is a gather
    K = cols.shape[1]
    mask = torch.arange(K, device=cols.device)[None, :] < cnt[:, None]
    cols = torch.where(mask, cols, torch.zeros_like(cols))
    vals = torch.where(mask, vals, torch.zeros_like(vals))
    for t in (cols, vals, cnt):
        dist.all_reduce(t, group=group)
    return cols, vals, cnt


__all__ = [
    "row_bounds_by_nnz", "shard_csr", "allgather_rows", "deal_by_cost", "ShardedImplicitMFTrainer",
    "sharded_knn_build_topk", "ALSTrainerBase",
]  # fmt: skip
