"""
Drop-in surface of the reference's native accelerator module for the two hot
paths — the same function names, argument meaning and error behaviour as
``lenskit._accel.als`` / ``lenskit._accel.knn`` (stubs
``src/lenskit/_accel/als.pyi``, ``knn.pyi``) and the ``AccelTask`` protocol of
``src/lenskit/parallel/_task.py:60-100`` — executed on the B200 through the C
ABI.  A maintainer rebinding ``from lenskit._accel import als`` /
``from lenskit import _accel`` to this module gets the GPU engine under the
unchanged Python components.

Host arrays in, host arrays out: ``this`` is mutated in place like the NumPy
borrow at ``implicit.rs:57-58``; device copies of the CSR operands are cached
per matrix object so that the ten half-epochs of a training run upload it once.
"""

from __future__ import annotations

import threading
from types import SimpleNamespace
from typing import Any, Callable, Generic, TypeVar

import numpy as np
import torch

from . import _lib, engine
from .data import InteractionCSR

R = TypeVar("R")
UPDATE_INTERVAL = 0.2  # parallel/_task.py:22


class AccelTask(Generic[R]):
    """
    Mirror of the Rust ``AccelTask`` pyclass (``src/accel/tasks/mod.rs:33-106``):
    ``invoke`` once, ``cancel`` and ``current_progress`` from another thread.
    """

    def __init__(self, fn: Callable[["AccelTask"], R], total: int | None = None):
        self._fn = fn
        self._invoked = False
        self._cancelled = threading.Event()
        self._progress = 0
        self._total = total

    def invoke(self, *, pool: Any = None) -> R:
        if self._invoked:
            raise RuntimeError("task already invoked")  # tasks/mod.rs:64-70
        self._invoked = True
        if self._cancelled.is_set():
            raise RuntimeError("task cancelled")
        res = self._fn(self)
        if self._total is not None:
            self._progress = self._total
        return res

    def cancel(self) -> None:
        self._cancelled.set()

    def current_progress(self) -> int | tuple[int, int] | None:
        if self._total is None:
            return None
        return (self._progress, self._total)


def run_accel_task(task: AccelTask[R], *, progress: Any = None) -> R:
    """``lenskit.parallel.run_accel_task`` (parallel/_task.py:25-57)."""
    box: dict[str, Any] = {}
    done = threading.Event()

    def body():
        try:
            box["v"] = task.invoke(pool=None)
        except BaseException as e:  # noqa: BLE001
            box["e"] = e
        done.set()

    th = threading.Thread(target=body, name="AccelTask", daemon=False)
    th.start()
    try:
        while not done.wait(UPDATE_INTERVAL if progress is not None else None):
            cp = task.current_progress()
            if progress is not None and isinstance(cp, tuple):
                progress.update(completed=cp[0], total=cp[1])
    except KeyboardInterrupt:
        task.cancel()
        raise
    th.join()
    if "e" in box:
        raise RuntimeError("accelerator task failed with exception") from box["e"]
    return box["v"]


# ---------------------------------------------------------------------------
# operand conversion / caching
# ---------------------------------------------------------------------------


def as_host_csr(m: Any) -> InteractionCSR:
    """Accept an ``InteractionCSR``, a SciPy CSR, or a ``SparseRowArray``-like Arrow array."""
    if isinstance(m, InteractionCSR):
        return m
    if hasattr(m, "offsets") and hasattr(m, "indices") and hasattr(m, "values"):  # SparseRowArray
        off = m.offsets.to_numpy(zero_copy_only=False)
        idx = m.indices.to_numpy(zero_copy_only=False)
        val = m.values.to_numpy(zero_copy_only=False)
        ncol = getattr(m, "dimension", int(idx.max()) + 1 if len(idx) else 0)
        return InteractionCSR(
            np.ascontiguousarray(off), np.ascontiguousarray(idx, dtype=np.int32),
            np.ascontiguousarray(val, dtype=np.float32), (len(off) - 1, int(ncol)),
        )  # fmt: skip
    if hasattr(m, "indptr") and hasattr(m, "data"):
        return InteractionCSR.from_scipy(m)
    raise TypeError(f"cannot interpret {type(m).__name__} as a CSR matrix")  # csr.rs:160-195


# device copies and plans of the host matrices the caller keeps passing in (one per epoch half):
# keyed by object identity, a few entries, oldest evicted first
_cache_keep: dict[tuple[int, str], tuple[Any, Any]] = {}
_CACHE_LIMIT = 8
# The cached plans / scorer states carry mutable device state (status word, work counter, Σ‖Δ‖²,
# contribution pool): calls into this mirror are serialised per process.  The reference's score_*
# are re-entrant on CPU threads; here concurrent callers queue on the one GPU stream anyway.
_engine_lock = threading.RLock()


def _cached(obj: Any, tag: str, make: Callable[[], Any]) -> Any:
    key = (id(obj), tag)
    with _engine_lock:
        ent = _cache_keep.get(key)
        if ent is not None and ent[0] is obj:
            return ent[1]
        val = make()
        if len(_cache_keep) >= _CACHE_LIMIT:
            _cache_keep.pop(next(iter(_cache_keep)))
        _cache_keep[key] = (obj, val)
        return val


def clear_cache() -> None:
    _cache_keep.clear()


def _check_factor(name: str, a: Any, writable: bool = False) -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.ndim != 2:
        raise TypeError(f"{name} must be a 2-D float32 ndarray")
    if not a.flags.c_contiguous:
        raise TypeError(f"{name} must be C-contiguous")
    if writable and not a.flags.writeable:
        raise TypeError(f"{name} must be writable")
    return a


def _als_task(mode: int, matrix: Any, this: np.ndarray, other: np.ndarray, otor, reg: float) -> AccelTask[float]:
    this = _check_factor("this", this, writable=True)
    other = _check_factor("other", other)
    if np.shares_memory(this, other):
        raise TypeError("this and other alias")  # numpy borrow check, implicit.rs:57-64
    n_rows, k = this.shape
    csr = as_host_csr(matrix)
    if csr.shape[0] != n_rows or csr.shape[1] != other.shape[0] or other.shape[1] != k:
        raise ValueError("matrix / factor shapes do not agree")
    if otor is not None:
        otor = _check_factor("otor", otor)
        if otor.shape != (k, k):
            raise ValueError("otor must be k x k")

    def run(task: AccelTask) -> float:
        with _engine_lock:
            return run_locked(task)

    def run_locked(task: AccelTask) -> float:
        dev = _lib.require_device()
        dm = _cached(matrix, "csr", lambda: engine.DeviceCSR.from_host(csr, dev))
        plan = _cached(matrix, f"plan{k}", lambda: engine.ALSHalfPlan.create(dm, k))
        d_this = torch.from_numpy(this).to(dev)
        d_other = torch.from_numpy(other).to(dev)
        d_otor = torch.from_numpy(otor).to(dev) if otor is not None else None
        plan.sqdelta.zero_()
        plan.status.zero_()
        engine.als_half_epoch(plan, mode, d_this, d_other, otor=d_otor, reg=reg)
        this[...] = d_this.cpu().numpy()
        st = int(plan.status.item())
        if st != 0:
            # solve.rs:99-105 → implicit.rs:79
            raise RuntimeError(f"ALS solve error: array minor of row {st - 1} is not positive")
        return float(np.sqrt(plan.sqdelta.item()))

    return AccelTask(run, total=n_rows)


def train_implicit_matrix(matrix: Any, this: np.ndarray, other: np.ndarray, otor: np.ndarray) -> AccelTask[float]:
    """``_accel.als.train_implicit_matrix`` (src/accel/als/implicit.rs:35-53)."""
    return _als_task(_lib.LK_ALS_IMPLICIT, matrix, this, other, otor, 0.0)


def train_explicit_matrix(matrix: Any, this: np.ndarray, other: np.ndarray, reg: float) -> AccelTask[float]:
    """``_accel.als.train_explicit_matrix`` (src/accel/als/explicit.rs:35-52)."""
    return _als_task(_lib.LK_ALS_EXPLICIT, matrix, this, other, None, float(reg))


def compute_similarities(
    ui_ratings: Any, iu_ratings: Any, shape: tuple[int, int], min_sim: float, save_nbrs: int | None
) -> AccelTask[list[InteractionCSR]]:
    """
    ``_accel.knn.compute_similarities`` (src/accel/knn/item_train.rs:32-93).  The
    reference returns a list of LargeList chunks in row order; this returns a
    one-element list holding the whole matrix (int64 offsets).
    """
    ui = as_host_csr(ui_ratings)
    iu = as_host_csr(iu_ratings)
    nu, ni = shape
    if ui.shape != (nu, ni) or iu.shape != (ni, nu):  # asserts at item_train.rs:51-54
        raise AssertionError("matrix shapes do not match `shape`")

    def run(task: AccelTask) -> list[InteractionCSR]:
        with _engine_lock:
            return run_locked(task)

    def run_locked(task: AccelTask) -> list[InteractionCSR]:
        dev = _lib.require_device()
        d_ui = _cached(ui_ratings, "csr", lambda: engine.DeviceCSR.from_host(ui, dev))
        d_iu = _cached(iu_ratings, "csr", lambda: engine.DeviceCSR.from_host(iu, dev))
        plan = engine.KnnBuildPlan.create(d_ui, d_iu)
        if save_nbrs is not None and save_nbrs > 0:
            cols, vals, cnt = plan.build_topk(min_sim, int(save_nbrs))
            indptr, c, v = engine.topk_rows_to_csr(cols, vals, cnt)
        else:
            indptr, c, v = plan.build_unbounded(min_sim)
        out = InteractionCSR(indptr.cpu().numpy(), c.cpu().numpy(), v.cpu().numpy(), (ni, ni))
        return [out]

    return AccelTask(run, total=ni)


def _null_to_neg(a: Any) -> np.ndarray:
    """Int32 item numbers with Arrow nulls (or masked entries) as -1."""
    if hasattr(a, "to_numpy") and hasattr(a, "null_count"):  # pyarrow
        import pyarrow.compute as pc

        return pc.fill_null(a, -1).to_numpy(zero_copy_only=False).astype(np.int32)
    return np.asarray(a, dtype=np.int32)


def _to_f32(a: Any) -> np.ndarray:
    if hasattr(a, "to_numpy") and hasattr(a, "null_count"):
        import pyarrow.compute as pc

        return pc.fill_null(a, 0.0).to_numpy(zero_copy_only=False).astype(np.float32)
    return np.asarray(a, dtype=np.float32)


def _score(sims: Any, ref_items, ref_rates, tgt_items, max_nbrs: int, min_nbrs: int):
    import pyarrow as pa

    s = as_host_csr(sims)
    if s.shape[0] != s.shape[1]:
        raise AssertionError("similarity matrix must be square")  # item_score.rs:116
    dev = _lib.require_device()
    st = _cached(
        sims, "score", lambda: engine.KnnScorerState.create(s.shape[0], s.indptr, s.indices, s.values, dev)
    )
    ri = _null_to_neg(ref_items)
    ti = _null_to_neg(tgt_items)
    rv = None if ref_rates is None else _to_f32(ref_rates)
    i64 = torch.int64
    # (KnnScorerState.score holds the state's own lock around its mutable device scratch)
    scores, counts = st.score(
        torch.tensor([0, len(ri)], dtype=i64, device=dev),
        torch.from_numpy(ri).to(dev),
        None if rv is None else torch.from_numpy(rv).to(dev),
        torch.tensor([0, len(ti)], dtype=i64, device=dev),
        torch.from_numpy(ti).to(dev),
        max_nbrs,
        min_nbrs,
    )
    sc = scores.cpu().numpy()
    ct = counts.cpu().numpy()
    return pa.array(sc, mask=np.isnan(sc)), pa.array(ct, mask=ct < 0)


def score_explicit(sims, ref_items, ref_rates, tgt_items, max_nbrs: int, min_nbrs: int):
    """``_accel.knn.score_explicit`` (src/accel/knn/item_score.rs:22-68)."""
    return _score(sims, ref_items, ref_rates, tgt_items, max_nbrs, min_nbrs)


def score_implicit(sims, ref_items, tgt_items, max_nbrs: int, min_nbrs: int):
    """``_accel.knn.score_implicit`` (src/accel/knn/item_score.rs:71-111)."""
    return _score(sims, ref_items, None, tgt_items, max_nbrs, min_nbrs)


def argtopn(scores: Any, n: int) -> np.ndarray:
    """
    ``_accel.data.argtopn`` (src/accel/data/sorting.rs:131-170): indices of the ``n`` largest non-NaN
    (and non-null) scores, in the reference's order.  One vector is one thread's work for
    ``lk_topn_columns`` — this entry point exists for interface parity; batches go through
    ``argtopn_batch`` / ``ALSBase.recommend_batch``.
    """
    s = _to_f32(scores)
    return argtopn_batch(s[None, :], n)[0]


def argtopn_batch(scores: Any, n: int) -> list[np.ndarray]:
    """``argtopn`` for every row of a host matrix ``scores`` [n_vectors, n_items] in one launch."""
    dev = _lib.require_device()
    s = np.ascontiguousarray(scores, dtype=np.float32)
    if s.ndim != 2:
        raise TypeError("scores must be 2-dimensional")
    n = int(min(n, s.shape[1]))
    if n <= 0 or s.shape[0] == 0:
        return [np.empty(0, dtype=np.int32) for _ in range(s.shape[0])]
    d = torch.from_numpy(s).to(dev).T.contiguous()  # item-major: the kernel scans a column per thread
    idx, _val, cnt = engine.topn_columns(d, n, with_values=False)
    idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
    return [idx[b, : cnt[b]].copy() for b in range(s.shape[0])]


#: ``from lkpy_b200.accel import als, knn`` mirrors ``from lenskit._accel import als, knn``
data = SimpleNamespace(argtopn=argtopn, argtopn_batch=argtopn_batch)
als = SimpleNamespace(
    train_implicit_matrix=train_implicit_matrix, train_explicit_matrix=train_explicit_matrix
)
knn = SimpleNamespace(
    compute_similarities=compute_similarities, score_explicit=score_explicit, score_implicit=score_implicit
)
