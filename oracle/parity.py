"""
Parity checks at the benchmark's own shapes — TEST INFRASTRUCTURE (oracle side).

``bench.py`` calls these after its timed loops (never inside them) and ``tests/test_scale_parity.py``
runs the same checks under ``-m gpu``: the engine's output for a *sample* of rows is compared with
the oracle run on exactly those rows, so that the ML-25M-shaped configuration the numbers are quoted
on (80 k-nonzero rows split in 20 parts, empty users, rows shorter than k, the three-half kNN
geometry, tie groups at scale) is checked, not only the small shapes the oracle can do in full.

Everything here takes and returns host (NumPy) data; the callers own the device side.
"""

from __future__ import annotations

import numpy as np

import oracle

from lkpy_b200.data import InteractionCSR


# ---------------------------------------------------------------------------
# row samples
# ---------------------------------------------------------------------------


def sample_als_rows(indptr: np.ndarray, k: int, chunk_nnz: int, n_random: int = 1500, seed: int = 0,
                    max_nnz: int | None = None, n_longest: int = 16, n_split: int = 48) -> np.ndarray:
    """
    Rows of one half-step to check: the longest rows (every row that the plan splits into parts is a
    candidate; the ``n_longest`` longest always), up to 64 empty rows, 400 rows shorter than ``k``, and
    ``n_random`` others — sorted, unique.  ``max_nnz`` bounds the length of a sampled row (the scalar f64
    oracle costs nnz * k^2 per row: at 100 M interactions and k = 128 the hottest rows would take minutes).
    """
    rng = np.random.default_rng(seed)
    n = np.diff(np.asarray(indptr, dtype=np.int64))
    ok = np.ones(len(n), dtype=bool) if max_nnz is None else n <= max_nnz
    order = np.argsort(-np.where(ok, n, -1), kind="stable")
    picks = [order[:n_longest]]
    split = np.flatnonzero((n > chunk_nnz) & ok)
    if len(split):
        picks.append(rng.choice(split, min(len(split), n_split), replace=False))
    empty = np.flatnonzero(n == 0)
    picks.append(empty[:64])
    short = np.flatnonzero((n > 0) & (n < k))
    if len(short):
        picks.append(rng.choice(short, min(len(short), 400), replace=False))
    cand = np.flatnonzero(ok)
    picks.append(rng.choice(cand, min(len(cand), n_random), replace=False))
    return np.unique(np.concatenate(picks)).astype(np.int64)


def sub_csr(csr: InteractionCSR, rows: np.ndarray) -> InteractionCSR:
    """The CSR restricted to ``rows`` (same columns)."""
    ip = np.asarray(csr.indptr, dtype=np.int64)
    lens = ip[rows + 1] - ip[rows]
    out_ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.arange(ip[r], ip[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
    return InteractionCSR(out_ip, csr.indices[idx], csr.values[idx], (len(rows), csr.shape[1]))


# ---------------------------------------------------------------------------
# ALS
# ---------------------------------------------------------------------------


def _otor64(other: np.ndarray, reg: float) -> np.ndarray:
    """OᵀO + reg·I accumulated in f64: the oracle's scalar loop where it finishes in seconds, a BLAS dgemm
    on the f64 copy of the same table beyond that (1 M x 128 rows: the two agree to ~1e-15 relative)."""
    n, k = other.shape
    if n * k * k <= 4e9:
        return oracle.otor(other, reg)[1]
    o = other.astype(np.float64)
    return o.T @ o + np.eye(k) * float(reg)


def check_als_half(
    mode: str, csr: InteractionCSR, rows: np.ndarray, this_old_rows: np.ndarray, other: np.ndarray,
    got_rows: np.ndarray, reg: float, bf16: bool, tol: float = 1e-4,
) -> dict:
    """
    ``got_rows`` = the engine's new ``this[rows]`` after one half-step from (``this_old``, ``other``);
    compared with the f64 oracle on the same rows (``bf16``: the oracle rounds the gathered rows to
    bf16 exactly as the engine's stored copy does — DESIGN.md §2).  Reports the relative Frobenius
    error over the sample (the north-star criterion), the worst row, and for bf16 also
    the distance to the *unrounded* f32-input oracle (reported, not gated: bf16 storage costs
    ~4e-3).
    """
    sub = sub_csr(csr, rows)
    k = other.shape[1]
    o_in = oracle.bf16_round(other) if bf16 else other
    if mode == "implicit":
        o64 = _otor64(o_in, reg)
        ref, _ = oracle.als_half_f64("implicit", sub, this_old_rows, other, otor_mat=o64, bf16_other=bf16)
    else:
        ref, _ = oracle.als_half_f64("explicit", sub, this_old_rows, other, reg=reg, bf16_other=bf16)
    got = np.asarray(got_rows, dtype=np.float64)
    err = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
    rn = np.linalg.norm(ref, axis=1)
    row_err = np.linalg.norm(got - ref, axis=1) / np.maximum(rn, 1e-30)
    n = np.diff(np.asarray(csr.indptr, dtype=np.int64))[rows]
    empty = n == 0
    out = {
        "rows": int(len(rows)),
        "longest_row_nnz": int(n.max()) if len(n) else 0,
        "empty_rows": int(empty.sum()),
        "rows_shorter_than_k": int(((n > 0) & (n < k)).sum()),
        "rel_fro_vs_f64_oracle": err,
        "row_rel_err_p99": float(np.quantile(row_err[~empty], 0.99)) if (~empty).any() else 0.0,
        "row_rel_err_max": float(row_err[~empty].max()) if (~empty).any() else 0.0,
        "empty_rows_zero": bool(np.all(got[empty] == 0.0)),
        "finite": bool(np.isfinite(got).all()),
        "tol": tol,
    }
    if bf16:
        if mode == "implicit":
            refu, _ = oracle.als_half_f64("implicit", sub, this_old_rows, other, otor_mat=_otor64(other, reg))
        else:
            refu, _ = oracle.als_half_f64("explicit", sub, this_old_rows, other, reg=reg)
        out["rel_fro_vs_unrounded_f64_oracle"] = float(np.linalg.norm(got - refu) / max(np.linalg.norm(refu), 1e-300))
    out["ok"] = bool(err <= tol and out["empty_rows_zero"] and out["finite"])
    return out


# ---------------------------------------------------------------------------
# item-kNN build
# ---------------------------------------------------------------------------


def sample_knn_rows(cost: np.ndarray, item_nnz: np.ndarray, n_random: int = 400, seed: int = 0) -> np.ndarray:
    """Item rows to check: the 8 most expensive, the 92 items with the fewest ratings (identical
    normalised vectors → tie groups at the cut), and ``n_random`` others."""
    rng = np.random.default_rng(seed)
    cost = np.asarray(cost)
    picks = [np.argsort(-cost, kind="stable")[:8], np.argsort(item_nnz, kind="stable")[:92]]
    picks.append(rng.choice(len(cost), min(len(cost), n_random), replace=False))
    return np.unique(np.concatenate(picks)).astype(np.int64)


def check_knn_rows(ui: InteractionCSR, iu: InteractionCSR, rows: np.ndarray, got_indptr: np.ndarray,
                   got_cols: np.ndarray, got_vals: np.ndarray, min_sim: float, save_nbrs: int | None) -> dict:
    """
    Bit-exact comparison (indices *and* value bits) of the engine's similarity rows ``rows`` with the
    oracle's ``sim_row`` (item_train.rs:95-152) on the same inputs.  ``got_*`` is the engine's whole
    CSR result (host arrays, int64 offsets).
    """
    bad_rows, ties = [], 0
    total = 0
    ref_all = oracle.knn_build(ui, iu, min_sim, save_nbrs, row_list=np.asarray(rows, dtype=np.int64))
    for t, r in enumerate(rows):
        ra, rb = int(ref_all.indptr[t]), int(ref_all.indptr[t + 1])
        rc, rv = ref_all.indices[ra:rb], ref_all.data[ra:rb]
        a, b = int(got_indptr[r]), int(got_indptr[r + 1])
        gc, gv = got_cols[a:b], got_vals[a:b]
        total += len(rc)
        if save_nbrs and len(rv) == save_nbrs:
            # a tie at the cut: more than one candidate shares the smallest kept similarity
            ties += int(np.sum(rv == rv.min()) > 1)
        same = (
            len(gc) == len(rc)
            and np.array_equal(gc, rc)
            and np.array_equal(np.ascontiguousarray(gv).view(np.int32), np.ascontiguousarray(rv).view(np.int32))
        )
        if not same:
            bad_rows.append(int(r))
    return {
        "rows": int(len(rows)), "neighbours_compared": int(total), "rows_with_tied_minimum": int(ties),
        "mismatched_rows": bad_rows[:16], "n_mismatched": len(bad_rows), "ok": not bad_rows,
    }  # fmt: skip


def checksum(*arrays: np.ndarray) -> int:
    """Order-sensitive 64-bit checksum of the raw bits of host arrays (cross-rank equality)."""
    import zlib

    h = 0
    for a in arrays:
        b = np.ascontiguousarray(a).view(np.uint8)
        h = (zlib.crc32(b, h & 0xFFFFFFFF) | (zlib.adler32(b, (h >> 32) & 0xFFFFFFFF) << 32)) & 0xFFFFFFFFFFFFFFFF
    return int(h)
