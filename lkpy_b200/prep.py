"""
Interaction-matrix preparation on the device: the O(nnz) host stages in front of
the two hot paths — CSR / CSC construction and transposition (``ALSTrainerBase.__init__``,
``src/lenskit/als/_common.py:216-219``; ``src/accel/data/transpose.rs:19-108``) and the item-kNN
centring + normalisation (``src/lenskit/knn/item.py:202-228``) — done in HBM so that a training call
uploads the COO triplets once and nothing else crosses PCIe.

The index plumbing (stable sorts, bincounts, prefix sums) uses torch's device primitives; the
arithmetic whose bits matter — the kNN column statistics — is ``lk_knn_prep_columns`` (``prep.cu``),
which reproduces NumPy's pairwise summation so that the f32 inputs of the bit-exact build are the
same bits SciPy produces (``tests/test_prep_gpu.py``).

``synth_interactions_device`` is the ML-25M-shaped generator of SURVEY.md §8d on the device, for the
scale-out configurations (100 M / 1 B interactions), where the NumPy generator would spend minutes
of host time per rank.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .engine import DeviceCSR


@dataclass
class DeviceInteractions:
    """COO interactions resident on the device, sorted by (user, item), deduplicated — what
    ``lkpy_b200.data.Interactions`` is on the host.  The ALS trainers and ``ItemKNNScorer.train`` accept
    it directly (no host round trip: the scale-out configurations are generated in HBM)."""

    users: torch.Tensor  # int32 [nnz]
    items: torch.Tensor  # int32 [nnz]
    ratings: torch.Tensor  # float32 [nnz]
    n_users: int
    n_items: int

    @property
    def nnz(self) -> int:
        return int(self.users.numel())


def _indptr(keys: torch.Tensor, n: int) -> torch.Tensor:
    """int32 CSR offsets of sorted integer keys in [0, n)."""
    counts = torch.bincount(keys, minlength=n)
    out = torch.zeros(n + 1, dtype=torch.int64, device=keys.device)
    torch.cumsum(counts, 0, out=out[1:])
    if int(out[-1].item()) >= 2**31:
        raise _lib.EngineError("matrices with nnz >= 2**31 need row sharding first")
    return out.to(torch.int32)


def coo_to_csr_pair(
    users: torch.Tensor, items: torch.Tensor, values: torch.Tensor, n_users: int, n_items: int
) -> tuple[DeviceCSR, DeviceCSR, torch.Tensor]:
    """
    Both orientations of a rating matrix given as COO triplets sorted by (user, item) on the device:
    R (users x items) and Rᵀ (items x users, users ascending inside a row — the role of
    ``SparseRowArray.from_scipy(coo.T)`` / ``transpose.rs``).  Returns (R, Rᵀ, perm) where
    ``values[perm]`` is the value array of Rᵀ.
    """
    u64, i64 = users.long(), items.long()
    ui_ptr = _indptr(u64, n_users)
    _sorted_items, perm = torch.sort(i64, stable=True)  # users stay ascending inside an item
    iu_ptr = _indptr(i64, n_items)
    h = lambda t: t.cpu().numpy()  # noqa: E731 — host copy of the offsets for the planners
    ui = DeviceCSR(ui_ptr, items.to(torch.int32).contiguous(), values.contiguous(), (n_users, n_items), h(ui_ptr))
    iu = DeviceCSR(iu_ptr, users[perm].to(torch.int32).contiguous(), values[perm].contiguous(), (n_items, n_users), h(iu_ptr))
    return ui, iu, perm


def knn_item_matrices_device(
    users: torch.Tensor, items: torch.Tensor, ratings: torch.Tensor | None, n_users: int, n_items: int, explicit: bool
) -> tuple[DeviceCSR, DeviceCSR, torch.Tensor | None]:
    """
    ``lkpy_b200.data.knn_item_matrices`` (the reference's ``_center_ratings`` + ``_normalize_rows``,
    ``knn/item.py:202-228``) on the device, bit-identical: returns (UI, IU, item means or None) with the
    normalised f32 values.  ``users``/``items``/``ratings`` are device COO triplets sorted by (user, item);
    implicit feedback (``explicit=False``) uses a value of 1 for every interaction.
    """
    dev = users.device
    vals = ratings.to(torch.float32) if explicit else torch.ones(users.numel(), dtype=torch.float32, device=dev)
    ui, iu, perm = coo_to_csr_pair(users, items, vals, n_users, n_items)
    means = torch.zeros(n_items, dtype=torch.float32, device=dev) if explicit else None
    out = torch.empty_like(iu.values)
    check(
        lib().lk_knn_prep_columns(ptr(iu.indptr), ptr(iu.values), n_items, 1 if explicit else 0, ptr(means), ptr(out), stream_ptr()),
        "lk_knn_prep_columns",
    )
    iu.values = out
    back = torch.empty_like(out)
    back[perm] = out  # the same values in (user, item) order
    ui.values = back
    return ui, iu, means


def synth_interactions_device(
    n_users: int, n_items: int, nnz: int, seed: int = 20260924, device=None
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """
    The synthetic generator of SURVEY.md §8d (``lkpy_b200.data.synth_interactions``) on the device:
    user weights LogNormal(0, 1.25), item weights (rank+20)^-1.05 with a seeded rank→id permutation,
    ⌈1.25·nnz⌉ i.i.d. pairs drawn from the two categoricals, de-duplicated, shuffled, cut to ``nnz``,
    sorted by (user, item); ratings i.i.d. from the ML-like pmf.  Same construction and distribution
    as the NumPy generator, different random streams (torch's Philox): the two do not produce the same
    matrix.  Returns (users int32, items int32, ratings f32), all on ``device``.
    """
    from .data import ML_RATING_PMF, ML_RATING_VALUES

    device = device or _lib.require_device()
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    # the two CDFs are accumulated on the host: a CUDA prefix sum of floats is not run-to-run (or
    # rank-to-rank) reproducible, and every rank of a sharded run must generate the same matrix
    def _cdf(w: torch.Tensor) -> torch.Tensor:
        h = w.cpu().numpy()
        c = np.cumsum(h / h.sum())
        c[-1] = 1.0
        return torch.from_numpy(c).to(device)

    ucdf = _cdf(torch.exp(torch.randn(n_users, generator=gen, device=device, dtype=torch.float64) * 1.25))
    icdf = _cdf((torch.arange(n_items, dtype=torch.float64) + 20.0) ** -1.05)
    perm = torch.randperm(n_items, generator=gen, device=device)
    want = int(np.ceil(1.25 * nnz))
    keys = torch.empty(0, dtype=torch.int64, device=device)
    for _round in range(8):
        chunks = []
        for lo in range(0, want, 1 << 27):  # bounded temporaries at 1 B interactions
            m = min(1 << 27, want - lo)
            us = torch.searchsorted(ucdf, torch.rand(m, generator=gen, device=device, dtype=torch.float64), right=True)
            ranks = torch.searchsorted(icdf, torch.rand(m, generator=gen, device=device, dtype=torch.float64), right=True)
            chunks.append(us.clamp_max(n_users - 1) * n_items + perm[ranks.clamp_max(n_items - 1)])
            del us, ranks
        new = torch.cat(chunks + ([keys] if keys.numel() else []))
        del chunks
        keys = torch.unique(new)  # sorted, distinct
        del new
        if keys.numel() >= nnz:
            break
    if keys.numel() < nnz:
        raise ValueError(f"could only draw {keys.numel()} distinct pairs of {nnz}")
    if keys.numel() > nnz:
        # "shuffle, keep the first nnz" = a uniform random subset: drop a uniform random sample of the surplus
        drop = torch.randperm(keys.numel(), generator=gen, device=device)[: keys.numel() - nnz]
        keep = torch.ones(keys.numel(), dtype=torch.bool, device=device)
        keep[drop] = False
        keys = keys[keep]  # still sorted by (user, item)
        del keep, drop
    users = (keys // n_items).to(torch.int32)
    items = (keys % n_items).to(torch.int32)
    del keys
    pcdf = torch.from_numpy(np.cumsum(ML_RATING_PMF / ML_RATING_PMF.sum())).to(device)
    rv = torch.searchsorted(pcdf, torch.rand(nnz, generator=gen, device=device, dtype=torch.float64), right=True)
    ratings = torch.tensor(ML_RATING_VALUES, device=device)[rv.clamp_max(len(ML_RATING_VALUES) - 1)]
    return users, items, ratings
