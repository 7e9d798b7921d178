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


def shard_device_csr(m: engine.DeviceCSR, lo: int, hi: int) -> engine.DeviceCSR:
    """Rows [lo, hi) of a device CSR as a device CSR of their own (offsets rebased)."""
    a, b = int(m.h_indptr[lo]), int(m.h_indptr[hi])
    hp = np.ascontiguousarray(m.h_indptr[lo : hi + 1] - m.h_indptr[lo]).astype(np.int32)
    return engine.DeviceCSR(
        (m.indptr[lo : hi + 1] - a).contiguous(), m.indices[a:b].contiguous(), m.values[a:b].contiguous(),
        (hi - lo, m.shape[1]), hp,
    )  # fmt: skip


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


def shared_pinned_tensor(name: str, shape: tuple[int, ...], dtype=torch.float32, create: bool = False) -> torch.Tensor:
    """
    A host tensor backed by ``/dev/shm/<name>`` and page-locked in this process: the ranks of one box
    map the same pages, so every rank can move *its own* rows of a host-resident model over *its own*
    PCIe link (``ShardedImplicitMFTrainer.train_epoch_e2e``).  The creator sizes the file; the others
    open it after a barrier.
    """
    n = int(np.prod(shape))
    path = f"/dev/shm/{name}"
    if create:
        with open(path, "wb") as f:
            f.truncate(n * torch.empty((), dtype=dtype).element_size())
    t = torch.from_file(path, shared=True, size=n, dtype=dtype).view(*shape)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
    if int(rc) != 0:
        raise _lib.EngineError(f"cudaHostRegister failed for {path} (code {int(rc)})")
    return t


def release_shared_pinned(t: torch.Tensor) -> None:
    """Undo the page-locking of ``shared_pinned_tensor`` BEFORE the tensor is dropped: a mapping that is
    unmapped while still registered leaves a stale registration behind, and a later host allocation that
    lands on the same addresses makes its first copy fail with "invalid argument"."""
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaHostUnregister(t.data_ptr())


class ShardedImplicitMFTrainer(ImplicitMFTrainer):
    """Row-sharded implicit ALS over the ranks of the default process group."""

    def __init__(self, scorer, data, options, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        super().__init__(scorer, data, options)
        ulo, uhi = self.u_slice
        ilo, ihi = self.i_slice
        # every rank must start from the same factors (rank 0's draw)
        dist.broadcast(self.d_users, 0, group=group)
        dist.broadcast(self.d_items, 0, group=group)
        self.user_peers = self.item_peers = None
        self._hu = self._hq = None
        if os.environ.get("LK_ALS_PEER_WRITES", "1") != "0":
            self._try_peer_tables()
        self._graph = None
        torch.cuda.empty_cache()

    def _make_plans(self, k: int) -> None:
        """Plans for this rank's row shards only (contiguous ranges balanced by nonzeros); the full
        matrices are replaced by the shards before any plan or workspace is sized."""
        self.u_bounds = row_bounds_by_nnz(self.ui.h_indptr, self.world)
        self.i_bounds = row_bounds_by_nnz(self.iu.h_indptr, self.world)
        ulo, uhi = int(self.u_bounds[self.rank]), int(self.u_bounds[self.rank + 1])
        ilo, ihi = int(self.i_bounds[self.rank]), int(self.i_bounds[self.rank + 1])
        self.ui = shard_device_csr(self.ui, ulo, uhi)
        self.iu = shard_device_csr(self.iu, ilo, ihi)
        self.u_slice = (ulo, uhi)
        self.i_slice = (ilo, ihi)
        super()._make_plans(k)

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
            self._hu = self._hq = None
            if self.rank == 0:
                print(f"[lkpy_b200] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather")

    def _exchange_barrier(self, hdl) -> None:
        """
        "Every rank's rows of this half have landed in every replica": a device-side barrier over the
        symmetric-memory signal pads — one tiny kernel on the stream, no NCCL collective, no host sync
        (round 1 used a scalar all-reduce here; the Σ‖Δ‖² partials are now reduced once per epoch, and
        only by the entry points that report them).
        """
        if hdl is not None and os.environ.get("LK_ALS_NCCL_BARRIER", "0") != "1":
            hdl.barrier(channel=0)
        else:
            dist.all_reduce(self._barrier_token(), group=self.group)

    def _barrier_token(self) -> torch.Tensor:
        t = self.__dict__.get("_token")
        if t is None:
            t = self.__dict__["_token"] = torch.zeros(1, device=self.device)
        return t

    def half_step(self, which: str, before_solve=None) -> torch.Tensor:
        """One sharded half-epoch ("user" | "item") including its exchange; returns this rank's Σ‖Δ‖² (device)."""
        if which == "user":
            plan, (lo, hi), this, other, obf = self.u_plan, self.u_slice, self.d_users, self.d_items, self.d_items_bf16
            reg, peers, hdl, bounds = self.config.user_reg, self.user_peers, self._hu, self.u_bounds
        else:
            plan, (lo, hi), this, other, obf = self.i_plan, self.i_slice, self.d_items, self.d_users, self.d_users_bf16
            reg, peers, hdl, bounds = self.config.item_reg, self.item_peers, self._hq, self.i_bounds
        if peers is not None:
            # peer-write path: rows land in every replica from inside the kernel
            d = self._half(plan, this[lo:hi], other, obf, reg, replicas=peers, replica_row0=lo, before_solve=before_solve)
            self._exchange_barrier(hdl)
        else:
            d = self._half(plan, this[lo:hi], other, obf, reg, before_solve=before_solve)
            allgather_rows(this, bounds, self.rank, self.world, self.group)
        return d

    def _epoch_body(self):
        self.u_plan.status.zero_()
        self.i_plan.status.zero_()
        du = self.half_step("user")
        di = self.half_step("item")
        return du, di

    def enable_graph(self) -> bool:
        """
        Capture the epoch (4 memsets, 6 kernels, 2 signal-pad barriers) in a CUDA graph: at 8 GPUs a
        half-epoch is ~0.3 ms of kernels and the Python-issued launches were a fixed ~0.2 ms per epoch.
        Only for the peer-write path (the NCCL all-gather fallback is not captured).
        """
        if self._graph is not None:
            return True
        if self.user_peers is None or self.kernel_events is not None:
            return False
        try:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._epoch_body()  # warm-up on the capture stream
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._epoch_body()
            self._graph = g
            self.epochs_trained += 1
        except Exception as e:  # noqa: BLE001
            self._graph = None
            if self.rank == 0:
                print(f"[lkpy_b200] CUDA graph capture of the epoch failed ({type(e).__name__}: {e}); eager launches")
        ok = torch.tensor([1 if self._graph is not None else 0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)  # all ranks or none: the barriers must pair up
        if int(ok.item()) == 0:
            self._graph = None
        return self._graph is not None

    def train_epoch_device(self):
        if self._graph is not None and self.kernel_events is None:
            self._graph.replay()
            du, di = self.u_plan.sqdelta, self.i_plan.sqdelta
        else:
            du, di = self._epoch_body()
        self.epochs_trained += 1
        return du, di

    def launches_per_epoch(self) -> int:
        return super().launches_per_epoch() + (2 if self.user_peers is not None else 0)  # + the signal-pad barriers

    def train_epoch_e2e(self, host_users: torch.Tensor, host_items: torch.Tensor) -> dict[str, float]:
        """
        The host-array epoch for the sharded trainer.  ``host_users`` / ``host_items`` are ONE host model
        shared by the ranks (``shared_pinned_tensor``): every rank uploads the rows it owns over its own
        PCIe link, the replicas are completed over NVLink, and every rank writes back only its own
        rows — P's rows underneath the item half (they are final after the user half).  No rank moves
        the whole model.  ``e2e_bytes`` holds this rank's (h2d, d2h) byte counts.
        """
        ulo, uhi = self.u_slice
        ilo, ihi = self.i_slice
        main = torch.cuda.current_stream()
        side = self._copy_stream()
        self.d_users[ulo:uhi].copy_(host_users[ulo:uhi], non_blocking=True)
        self.d_items[ilo:ihi].copy_(host_items[ilo:ihi], non_blocking=True)
        allgather_rows(self.d_items, self.i_bounds, self.rank, self.world, self.group)
        allgather_rows(self.d_users, self.u_bounds, self.rank, self.world, self.group)
        self.u_plan.status.zero_()
        self.i_plan.status.zero_()
        du = self.half_step("user")
        side.wait_stream(main)
        with torch.cuda.stream(side):
            host_users[ulo:uhi].copy_(self.d_users[ulo:uhi], non_blocking=True)  # under the item half
        di = self.half_step("item")
        self.epochs_trained += 1
        host_items[ilo:ihi].copy_(self.d_items[ilo:ihi], non_blocking=True)
        main.wait_stream(side)
        d = torch.cat([du, di])
        dist.all_reduce(d, group=self.group)  # Σ‖Δ‖² over the shards: the step's result
        k4 = self.d_users.shape[1] * 4
        nb = ((uhi - ulo) + (ihi - ilo)) * k4
        self.e2e_bytes = (nb, nb + 16)
        deltas = d.cpu()  # device->host read; synchronises (the shared model is complete after the next barrier)
        return {"deltaP": float(np.sqrt(deltas[0])), "deltaQ": float(np.sqrt(deltas[1]))}

    def train_epoch(self):
        du, di = self.train_epoch_device()
        d = torch.cat([du, di])
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
    """
    Item-sharded truncated build; every rank returns the full fixed-width result.  Items are dealt to
    the ranks in descending cost, snake order (``deal_by_cost`` on the device, no host sync); each
    rank builds the rows it owns (its hot items cut into column pieces sized for ``plan.world`` GPUs);
    ONE all-gather of the owned rows, packed as [cols | value bits | count], completes the result.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = plan.order.device
    n_items = plan.order.numel()
    pos = torch.arange(n_items, device=dev)
    lap, off = pos // world, pos % world
    owner = torch.where(lap % 2 == 0, off, world - 1 - off)
    order = plan.order[owner == rank].contiguous()
    cols, vals, cnt = plan.build_topk(min_sim, save_nbrs, order)
    K = cols.shape[1]
    # every rank owns ceil or floor(n_items / world) rows: pad the packed block to the ceiling
    m = -(-n_items // world)
    send = torch.zeros((m, 2 * K + 1), dtype=torch.int32, device=dev)
    o = order.long()
    send[: o.numel(), :K] = cols[o]
    send[: o.numel(), K : 2 * K] = vals[o].view(torch.int32)
    send[: o.numel(), 2 * K] = cnt[o]
    recv = torch.empty((world, m, 2 * K + 1), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv.view(world * m, 2 * K + 1), send, group=group)
    for r in range(world):
        if r == rank:
            continue
        rows = plan.order[owner == r].long()
        blk = recv[r, : rows.numel()]
        cols[rows] = blk[:, :K]
        vals[rows] = blk[:, K : 2 * K].view(torch.float32)
        cnt[rows] = blk[:, 2 * K]
    return cols, vals, cnt


__all__ = [
    "row_bounds_by_nnz", "shard_csr", "allgather_rows", "deal_by_cost", "ShardedImplicitMFTrainer",
    "sharded_knn_build_topk", "ALSTrainerBase", "shared_pinned_tensor", "release_shared_pinned",
]  # fmt: skip
