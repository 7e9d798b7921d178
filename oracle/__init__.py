"""
CPU oracle for the ALS / item-kNN hot paths — TEST INFRASTRUCTURE ONLY.

Thin ctypes binding of ``oracle/liblk_oracle.so`` (``oracle/lk_oracle.c``, a C
restatement of ``src/accel/als/*.rs`` and ``src/accel/knn/*.rs``).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; nothing under
``lkpy_b200/`` does.  See the header of ``lk_oracle.c`` for how it is pinned.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import scipy.sparse as sps

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "liblk_oracle.so"
_lib = None
_blas_limit = None

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


_FAST_PATH = _DIR / "liblk_cpu_fast.so"
_fast = None


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, no GPU needed)."""
    stale = force
    for so, src in ((_LIB_PATH, _DIR / "lk_oracle.c"), (_FAST_PATH, _DIR / "lk_cpu_fast.c")):
        stale = stale or not so.exists() or so.stat().st_mtime < src.stat().st_mtime
    if stale:
        subprocess.run(["make", "-C", str(_DIR), "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def fast_lib():
    """``liblk_cpu_fast.so``: the many-core restatement of the ALS half-epoch timed as the CPU baseline."""
    global _fast
    if _fast is None:
        build()
        L = C.CDLL(str(_FAST_PATH))
        L.lk_cpu_als_half_f32.argtypes = [
            C.c_int, _i64p, _i32p, _f32p, C.c_int64, C.c_int, _f32p, _f32p, C.c_void_p,
            C.c_float, C.c_int, C.POINTER(C.c_double),
        ]  # fmt: skip
        L.lk_cpu_als_half_f32.restype = C.c_int64
        _fast = L
    return _fast


def als_half_fast(
    mode: str, matrix, this: np.ndarray, other: np.ndarray, *, otor_mat: np.ndarray | None = None,
    reg: float = 0.0, threads: int = 0, inplace: bool = False,
) -> tuple[np.ndarray, float]:
    """
    The timed CPU baseline: same half-epoch as ``als_half`` through ``lk_cpu_fast.c`` (thread-private
    register-blocked Gram, vectorised Cholesky; scales on many-core hosts).  ``inplace`` skips the
    defensive copy of ``this`` (timing runs).
    """
    indptr, cols, vals = _csr_parts(matrix)
    this = np.ascontiguousarray(this, dtype=np.float32) if inplace else np.array(this, dtype=np.float32, order="C")
    other = np.ascontiguousarray(other, dtype=np.float32)
    n_rows, k = this.shape
    m = 0 if mode == "implicit" else 1
    op = None
    if m == 0:
        assert otor_mat is not None
        otor_mat = np.ascontiguousarray(otor_mat, dtype=np.float32)
        op = otor_mat.ctypes.data_as(C.c_void_p)
    sq = C.c_double(0.0)
    fail = fast_lib().lk_cpu_als_half_f32(m, indptr, cols, vals, n_rows, k, this, other, op, reg, threads, C.byref(sq))
    if fail:
        raise RuntimeError(f"ALS solve error: row {fail - 1} not positive definite")
    return this, float(np.sqrt(sq.value))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.lk_oracle_set_blas.argtypes = [C.c_void_p, C.c_void_p]
        L.lk_oracle_set_blas.restype = None
        L.lk_oracle_max_threads.restype = C.c_int
        L.lk_oracle_bf16_round.argtypes = [_f32p, _f32p, C.c_int64]
        L.lk_oracle_bf16_round.restype = None
        L.lk_oracle_posv_f32.argtypes = [_f32p, _f32p, C.c_int]
        L.lk_oracle_posv_f32.restype = C.c_int
        L.lk_oracle_otor.argtypes = [_f32p, C.c_int64, C.c_int, C.c_float, C.c_int, _f32p, _f64p]
        L.lk_oracle_otor.restype = None
        L.lk_oracle_als_half_f32.argtypes = [
            C.c_int, _i64p, _i32p, _f32p, C.c_int64, C.c_int, _f32p, _f32p, C.c_void_p,
            C.c_float, C.c_int, C.c_int, C.POINTER(C.c_double),
        ]  # fmt: skip
        L.lk_oracle_als_half_f32.restype = C.c_int64
        L.lk_oracle_als_half_f64.argtypes = [
            C.c_int, _i64p, _i32p, _f32p, C.c_int64, C.c_int, _f32p, _f64p, _f32p, C.c_void_p,
            C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double),
        ]  # fmt: skip
        L.lk_oracle_als_half_f64.restype = C.c_int64
        L.lk_oracle_knn_build.argtypes = [
            _i64p, _i32p, _f32p, _i64p, _i32p, _f32p, C.c_int64, C.c_int64, C.c_float,
            C.c_int64, C.c_int64, C.c_int64, C.c_int,
        ]  # fmt: skip
        L.lk_oracle_knn_build.restype = C.c_void_p
        L.lk_oracle_knn_build_rows.argtypes = [
            _i64p, _i32p, _f32p, _i64p, _i32p, _f32p, C.c_int64, C.c_float, C.c_int64, _i64p, C.c_int64, C.c_int,
        ]  # fmt: skip
        L.lk_oracle_knn_build_rows.restype = C.c_void_p
        for nm, rt in (
            ("lk_oracle_csr_indptr", C.POINTER(C.c_int64)),
            ("lk_oracle_csr_cols", C.POINTER(C.c_int32)),
            ("lk_oracle_csr_vals", C.POINTER(C.c_float)),
            ("lk_oracle_csr_rows", C.c_int64),
        ):
            getattr(L, nm).argtypes = [C.c_void_p]
            getattr(L, nm).restype = rt
        L.lk_oracle_csr_free.argtypes = [C.c_void_p]
        L.lk_oracle_csr_free.restype = None
        L.lk_oracle_knn_score.argtypes = [
            _i64p, _i32p, _f32p, C.c_int64, _i32p, C.c_void_p, C.c_int64, _i32p, C.c_int64,
            C.c_int, C.c_int, _f32p, _i32p,
        ]  # fmt: skip
        L.lk_oracle_knn_score.restype = C.c_int
        L.lk_oracle_user_score.argtypes = [
            _i64p, _i32p, C.c_void_p, C.c_int64, C.c_int64, _i32p, _f32p, C.c_int64, _i32p, C.c_int64,
            C.c_int, C.c_int, _f32p, _i32p,
        ]  # fmt: skip
        L.lk_oracle_user_score.restype = C.c_int
        L.lk_oracle_argtopn_f32.argtypes = [_f32p, C.c_int64, C.c_int64, _i32p]
        L.lk_oracle_argtopn_f32.restype = C.c_int64
        _lib = L
    return _lib


def _capsule_ptr(module, name: str) -> int:
    cap = module.__pyx_capi__[name]
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap))


def use_scipy_blas(enable: bool = True) -> None:
    """
    Route the per-row solve through SciPy's bundled LAPACK ``sposv`` — the
    routine the reference resolves at ``src/accel/als/solve.rs:47-58`` via the
    same ``__pyx_capi__`` capsule — and the Gram through OpenBLAS ``sgemm``
    (the role of ``matrixmultiply`` at ``implicit.rs:112``).
    """
    if enable:
        from scipy.linalg import cython_blas, cython_lapack

        try:  # BLAS threads = 1 inside the row-parallel region
            import threadpoolctl

            global _blas_limit
            _blas_limit = threadpoolctl.threadpool_limits(limits=1, user_api="blas")
        except Exception:  # noqa: BLE001
            pass

        global _blas_thread_cap
        _blas_thread_cap = _openblas_caller_cap()
        lib().lk_oracle_set_blas(
            _capsule_ptr(cython_lapack, "sposv"), _capsule_ptr(cython_blas, "sgemm")
        )
    else:
        _blas_thread_cap = 0
        lib().lk_oracle_set_blas(None, None)


_blas_thread_cap = 0


def _openblas_caller_cap() -> int:
    """
    SciPy's OpenBLAS keeps a fixed table of 2*MAX_THREADS work buffers shared by its own server
    threads (one each) and every thread that calls into it; overflowing the table is a known crash
    ("Bad memory unallocation").  Keep the number of OpenMP threads that call sgemm/sposv
    concurrently safely below that: MAX_THREADS - 16 (48 for the MAX_THREADS=64 wheels).
    """
    cap = 48
    try:
        import re

        import threadpoolctl

        for info in threadpoolctl.threadpool_info():
            if info.get("internal_api") != "openblas" or "scipy.libs" not in info.get("filepath", ""):
                continue
            L = C.CDLL(info["filepath"])
            for name in ("scipy_openblas_get_config", "scipy_openblas_get_config64_", "openblas_get_config"):
                f = getattr(L, name, None)
                if f is None:
                    continue
                f.restype = C.c_char_p
                m = re.search(rb"MAX_THREADS=(\d+)", f() or b"")
                if m:
                    cap = max(1, int(m.group(1)) - 16)
                break
    except Exception:  # noqa: BLE001
        pass
    return cap


def blas_threads(threads: int = 0) -> int:
    """The thread count ``als_half`` really uses for ``threads`` while the BLAS path is enabled."""
    t = threads if threads > 0 else max_threads()
    return min(t, _blas_thread_cap) if _blas_thread_cap else t


def max_threads() -> int:
    return int(lib().lk_oracle_max_threads())


def bf16_round(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().lk_oracle_bf16_round(x.reshape(-1), out.reshape(-1), x.size)
    return out


def posv(A: np.ndarray, b: np.ndarray) -> tuple[np.ndarray, int]:
    A = np.array(A, dtype=np.float32, order="C")
    x = np.array(b, dtype=np.float32, order="C")
    info = lib().lk_oracle_posv_f32(A, x, len(x))
    return x, int(info)


def otor(other: np.ndarray, reg: float, bf16: bool = False) -> tuple[np.ndarray, np.ndarray]:
    """``_implicit_otor`` (als/_implicit.py:177-184): (f32-accumulated, f64-accumulated)."""
    other = np.ascontiguousarray(other, dtype=np.float32)
    n, k = other.shape
    o32 = np.empty((k, k), dtype=np.float32)
    o64 = np.empty((k, k), dtype=np.float64)
    lib().lk_oracle_otor(other, n, k, reg, int(bf16), o32, o64)
    return o32, o64


def _csr_parts(m):
    """Accept an ``InteractionCSR``-like object or a SciPy CSR."""
    if hasattr(m, "indptr") and hasattr(m, "indices"):
        indptr, cols = m.indptr, m.indices
        vals = m.values if hasattr(m, "values") else m.data
    else:
        raise TypeError("expected CSR")
    return (
        np.ascontiguousarray(indptr, dtype=np.int64),
        np.ascontiguousarray(cols, dtype=np.int32),
        np.ascontiguousarray(vals, dtype=np.float32),
    )


def als_half(
    mode: str,
    matrix,
    this: np.ndarray,
    other: np.ndarray,
    *,
    otor_mat: np.ndarray | None = None,
    reg: float = 0.0,
    bf16_other: bool = False,
    threads: int = 0,
) -> tuple[np.ndarray, float]:
    """
    One f32 ALS half-epoch (``train_implicit_matrix`` / ``train_explicit_matrix``).
    Returns (new ``this``, sqrt(sum ||delta||^2)); the input is not modified.
    """
    indptr, cols, vals = _csr_parts(matrix)
    this = np.array(this, dtype=np.float32, order="C")
    other = np.ascontiguousarray(other, dtype=np.float32)
    n_rows, k = this.shape
    m = 0 if mode == "implicit" else 1
    op = None
    if m == 0:
        assert otor_mat is not None
        otor_mat = np.ascontiguousarray(otor_mat, dtype=np.float32)
        op = otor_mat.ctypes.data_as(C.c_void_p)
    sq = C.c_double(0.0)
    if _blas_thread_cap and not bf16_other:
        threads = blas_threads(threads)
    fail = lib().lk_oracle_als_half_f32(
        m, indptr, cols, vals, n_rows, k, this, other, op, reg, int(bf16_other), threads, C.byref(sq)
    )
    if fail:
        raise RuntimeError(f"ALS solve error: row {fail - 1} not positive definite")
    return this, float(np.sqrt(sq.value))


def als_half_f64(
    mode: str,
    matrix,
    this: np.ndarray,
    other: np.ndarray,
    *,
    otor_mat: np.ndarray | None = None,
    reg: float = 0.0,
    bf16_other: bool = False,
    threads: int = 0,
) -> tuple[np.ndarray, float]:
    """f64 tolerance oracle of the same half-epoch; returns a float64 matrix."""
    indptr, cols, vals = _csr_parts(matrix)
    this = np.ascontiguousarray(this, dtype=np.float32)
    other = np.ascontiguousarray(other, dtype=np.float32)
    n_rows, k = this.shape
    out = np.empty((n_rows, k), dtype=np.float64)
    m = 0 if mode == "implicit" else 1
    op = None
    if m == 0:
        assert otor_mat is not None
        otor_mat = np.ascontiguousarray(otor_mat, dtype=np.float64)
        op = otor_mat.ctypes.data_as(C.c_void_p)
    sq = C.c_double(0.0)
    fail = lib().lk_oracle_als_half_f64(
        m, indptr, cols, vals, n_rows, k, this, out, other, op, reg, int(bf16_other), threads,
        C.byref(sq),
    )  # fmt: skip
    if fail:
        raise RuntimeError(f"ALS solve error: row {fail - 1} not positive definite")
    return out, float(np.sqrt(sq.value))


_parts_cache: dict = {}


def knn_operands(ui, iu):
    """The six contiguous arrays ``knn_build`` passes to C, converted once (int64 offsets) and kept
    for repeated calls on the same pair of matrices (row-by-row parity checks, timing repeats)."""
    key = (id(ui), id(iu))
    ent = _parts_cache.get(key)
    if ent is None or ent[0] is not ui or ent[1] is not iu:
        _parts_cache.clear()
        ent = (ui, iu, _csr_parts(ui), _csr_parts(iu))
        _parts_cache[key] = ent
    return ent[2], ent[3]


def knn_build(
    ui, iu, min_sim: float, save_nbrs: int | None, rows: tuple[int, int] | None = None, threads: int = 0,
    row_list: np.ndarray | None = None,
) -> sps.csr_array:
    """``compute_similarities`` (item_train.rs:32-152) → CSR with int64 offsets.  ``rows`` = a
    contiguous range of item rows, ``row_list`` = an explicit list (result rows in list order)."""
    (uip, uic, uiv), (iup, iuc, iuv) = knn_operands(ui, iu)
    n_users = len(uip) - 1
    n_items = len(iup) - 1
    L = lib()
    if row_list is not None:
        rl = np.ascontiguousarray(row_list, dtype=np.int64)
        h = L.lk_oracle_knn_build_rows(
            uip, uic, uiv, iup, iuc, iuv, n_items, np.float32(min_sim),
            int(save_nbrs) if save_nbrs else 0, rl, len(rl), threads,
        )  # fmt: skip
    else:
        rb, re = rows if rows is not None else (0, n_items)
        h = L.lk_oracle_knn_build(
            uip, uic, uiv, iup, iuc, iuv, n_users, n_items, np.float32(min_sim),
            int(save_nbrs) if save_nbrs else 0, rb, re, threads,
        )  # fmt: skip
    try:
        nr = L.lk_oracle_csr_rows(h)
        indptr = np.ctypeslib.as_array(L.lk_oracle_csr_indptr(h), shape=(nr + 1,)).copy()
        nnz = int(indptr[-1])
        if nnz:
            cols = np.ctypeslib.as_array(L.lk_oracle_csr_cols(h), shape=(nnz,)).copy()
            vals = np.ctypeslib.as_array(L.lk_oracle_csr_vals(h), shape=(nnz,)).copy()
        else:
            cols = np.empty(0, dtype=np.int32)
            vals = np.empty(0, dtype=np.float32)
    finally:
        L.lk_oracle_csr_free(h)
    return sps.csr_array((vals, cols, indptr), shape=(nr, n_items))


def knn_score(
    sims: sps.csr_array,
    ref_items: np.ndarray,
    ref_vals: np.ndarray | None,
    tgt_items: np.ndarray,
    max_nbrs: int,
    min_nbrs: int,
) -> tuple[np.ndarray, np.ndarray]:
    """``score_explicit`` / ``score_implicit`` (item_score.rs:22-111); NaN / -1 mark nulls."""
    indptr = np.ascontiguousarray(sims.indptr, dtype=np.int64)
    cols = np.ascontiguousarray(sims.indices, dtype=np.int32)
    vals = np.ascontiguousarray(sims.data, dtype=np.float32)
    ref_items = np.ascontiguousarray(ref_items, dtype=np.int32)
    tgt_items = np.ascontiguousarray(tgt_items, dtype=np.int32)
    rv = None
    if ref_vals is not None:
        ref_vals = np.ascontiguousarray(ref_vals, dtype=np.float32)
        rv = ref_vals.ctypes.data_as(C.c_void_p)
    scores = np.empty(len(tgt_items), dtype=np.float32)
    counts = np.empty(len(tgt_items), dtype=np.int32)
    rc = lib().lk_oracle_knn_score(
        indptr, cols, vals, sims.shape[0], ref_items, rv, len(ref_items), tgt_items,
        len(tgt_items), max_nbrs, min_nbrs, scores, counts,
    )  # fmt: skip
    if rc == 1:
        raise ValueError("similarity is null")
    if rc:
        raise IndexError("item index out of range")
    return scores, counts


def user_score(
    ratings, nbr_rows: np.ndarray, nbr_sims: np.ndarray, tgt_items: np.ndarray, max_nbrs: int, min_nbrs: int,
    explicit: bool = True,
) -> tuple[np.ndarray, np.ndarray]:
    """``user_score_items_explicit`` / ``_implicit`` (user_score.rs:21-98) on a users x items CSR of
    (centred) ratings; NaN / -1 mark nulls."""
    indptr, cols, vals = _csr_parts(ratings)
    n_users = len(indptr) - 1
    n_items = int(ratings.shape[1])
    nbr_rows = np.ascontiguousarray(nbr_rows, dtype=np.int32)
    nbr_sims = np.ascontiguousarray(nbr_sims, dtype=np.float32)
    tgt_items = np.ascontiguousarray(tgt_items, dtype=np.int32)
    scores = np.empty(len(tgt_items), dtype=np.float32)
    counts = np.empty(len(tgt_items), dtype=np.int32)
    rc = lib().lk_oracle_user_score(
        indptr, cols, vals.ctypes.data_as(C.c_void_p) if explicit else None, n_users, n_items, nbr_rows, nbr_sims,
        len(nbr_rows), tgt_items, len(tgt_items), max_nbrs, min_nbrs, scores, counts,
    )  # fmt: skip
    if rc == 1:
        raise ValueError("similarity is null")
    if rc:
        raise IndexError("index out of range")
    return scores, counts


def argtopn(scores: np.ndarray, n: int) -> np.ndarray:
    """
    ``lenskit._accel.data.argtopn`` (``src/accel/data/sorting.rs:131-170``): indices of the ``n``
    largest non-NaN scores in descending score order, selected and ordered by the reference's
    indirect min-heap (``src/accel/indirect/heap.rs``) — including which of several equal scores
    survive at the cut and the order equal scores come out in.
    """
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = int(min(max(n, 0), len(scores)))
    out = np.empty(max(n, 1), dtype=np.int32)
    cnt = lib().lk_oracle_argtopn_f32(scores, len(scores), n, out)
    return out[:cnt].copy()
