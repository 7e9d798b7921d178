"""
The argtopn oracle (``oracle.argtopn``: src/accel/data/sorting.rs:131-170 + src/accel/indirect/heap.rs)
against the reference's own unit-test vectors (heap.rs:104-162) and the properties its hypothesis
tests assert (tests/accel/test_argsort.py:60-200), and the GPU selection kernel
(``lk_topn_columns``) against the oracle — bit for bit, ties and NaNs included.
"""

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as nph

import oracle


def _check_properties(arr: np.ndarray, n: int, idx: np.ndarray) -> None:
    """test_topn_with_nans (tests/accel/test_argsort.py:119-146) for float32 input."""
    tgt_n = min(n, len(arr) - int(np.sum(np.isnan(arr))))
    assert len(idx) == tgt_n
    assert np.all(idx >= 0) and np.all(idx < max(len(arr), 1))
    assert len(set(idx.tolist())) == len(idx)
    items = arr[idx]
    assert not np.any(np.isnan(items))
    assert np.all(items[1:] <= items[:-1])  # descending
    if len(items):
        mask = np.ones(len(arr), dtype=np.bool_)
        mask[idx] = False
        mask[np.isnan(arr)] = False
        assert np.all(arr[mask] <= np.min(items))


def test_oracle_reference_heap_vectors():
    # heap.rs:118-162: two elements in either order, and 1..10 through a heap of 5
    assert oracle.argtopn(np.array([10, 20], np.float32), 5).tolist() == [1, 0]
    assert oracle.argtopn(np.array([20, 10], np.float32), 5).tolist() == [0, 1]
    assert oracle.argtopn(np.arange(1, 11, dtype=np.float32), 5).tolist() == [9, 8, 7, 6, 5]
    assert oracle.argtopn(np.array([10], np.float32), 5).tolist() == [0]
    assert oracle.argtopn(np.empty(0, np.float32), 5).tolist() == []
    assert oracle.argtopn(np.arange(5, dtype=np.float32), 0).tolist() == []


@settings(max_examples=300, deadline=None)
@given(
    nph.arrays(np.float32, nph.array_shapes(max_dims=1, max_side=400), elements={"allow_nan": True}),
    st.integers(min_value=0, max_value=500),
)
def test_oracle_topn_properties(arr, n):
    _check_properties(arr, n, oracle.argtopn(arr, n))


def test_oracle_topn_distinct_equals_argsort():
    rng = np.random.default_rng(3)
    a = rng.permutation(5000).astype(np.float32)
    for n in (1, 10, 100, 5000):
        assert np.array_equal(oracle.argtopn(a, n), np.argsort(-a, kind="stable")[:n].astype(np.int32))


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------


def _matrix(rng, n_items, n_vec, ties: bool, nan_frac: float) -> np.ndarray:
    s = rng.standard_normal((n_items, n_vec)).astype(np.float32)
    if ties:
        s = (np.round(s * 4) / 4).astype(np.float32)  # ~25 distinct values: ties everywhere
    if nan_frac:
        s[rng.random(s.shape) < nan_frac] = np.nan
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 16, 20, 100, 128, 300, 1024])
@pytest.mark.parametrize("ties", [False, True])
def test_gpu_topn_columns_bit_exact(n, ties):
    import torch

    from lkpy_b200 import _lib, engine

    dev = _lib.require_device()
    rng = np.random.default_rng(100 + n + ties)
    n_items, n_vec = (1500, 333) if n <= 300 else (2500, 70)
    s = _matrix(rng, n_items, n_vec, ties, 0.05)
    s[:, 5] = np.nan  # a vector with nothing to rank
    s[: n_items - 3, 6] = np.nan  # fewer valid scores than n
    idx, val, cnt = engine.topn_columns(torch.from_numpy(s).to(dev), n)
    idx, val, cnt = idx.cpu().numpy(), val.cpu().numpy(), cnt.cpu().numpy()
    for c in range(n_vec):
        ref = oracle.argtopn(s[:, c], n)
        assert cnt[c] == len(ref)
        assert np.array_equal(idx[c, : cnt[c]], ref), c
        assert np.all(idx[c, cnt[c] :] == -1)
        assert np.array_equal(val[c, : cnt[c]].view(np.int32), s[ref, c].view(np.int32))
        assert np.isnan(val[c, cnt[c] :]).all()


@pytest.mark.gpu
def test_gpu_topn_strided_view_and_api_mirror():
    import torch

    from lkpy_b200 import _lib, accel, engine

    dev = _lib.require_device()
    rng = np.random.default_rng(9)
    s = _matrix(rng, 800, 64, True, 0.1)
    d = torch.from_numpy(s).to(dev)
    # a column slice of a wider matrix: leading dimension != number of vectors
    idx, _v, cnt = engine.topn_columns(d[:, 10:40], 25, with_values=False)
    for c in range(30):
        ref = oracle.argtopn(s[:, 10 + c], 25)
        assert np.array_equal(idx[c, : int(cnt[c])].cpu().numpy(), ref)
    # `_accel.data.argtopn` mirror on one host vector, and the batch form on rows
    v = s[:, 3].copy()
    for n in (0, 1, 50, 800, 5000):
        got = accel.data.argtopn(v, n)
        assert np.array_equal(got, oracle.argtopn(v, n))
        _check_properties(v, min(n, len(v)), got)
    rows = accel.data.argtopn_batch(s.T[:8], 40)
    for r in range(8):
        assert np.array_equal(rows[r], oracle.argtopn(s[:, r], 40))
    # a nullable Arrow vector (what the kNN entry points return): nulls are unscored like NaNs
    # (sorting.rs:163-167), and the answer is an Int32Array like the reference's
    import pyarrow as pa

    masked = np.zeros(len(v), dtype=bool)
    masked[::5] = True
    got = accel.data.argtopn(pa.array(v, mask=masked), 25)
    assert isinstance(got, pa.Array) and got.type == pa.int32()
    assert got.to_numpy().tolist() == oracle.argtopn(np.where(masked, np.nan, v).astype(np.float32), 25).tolist()
    with pytest.raises(ValueError):
        engine.topn_columns(d, 0)
    with pytest.raises(ValueError):
        engine.topn_columns(d, 5000)


@pytest.mark.gpu
def test_gpu_als_recommend_batch(ml_small):
    """Batched fold-in + scoring + top-N against the per-query host path of the same scorer."""
    from lkpy_b200.als import BiasedMFScorer, ImplicitMFScorer
    from lkpy_b200.components import Dataset, ItemList, RecQuery, TrainingOptions

    ds = Dataset(ml_small)
    algo = ImplicitMFScorer(features=32, epochs=3)
    algo.train(ds, TrainingOptions(rng=7))
    all_items = ItemList(ml_small.item_ids)
    users = [3, 50, 400]
    queries = [RecQuery(user_id=ml_small.user_ids[u]) for u in users]
    queries += [RecQuery(user_id=None, query_items=ds.user_history(u)) for u in (10, 77)]
    queries += [RecQuery(user_id=-12345)]  # unknown, no history: nothing to recommend
    S = algo.score_matrix(queries).cpu().numpy()
    assert S.shape == (ml_small.n_items, len(queries))
    for b, q in enumerate(queries[:-1]):
        host = algo(q, all_items).scores()
        assert np.allclose(S[:, b], host, rtol=2e-4, atol=2e-5), b  # GEMM vs matvec summation order; fold-in solve
    assert np.isnan(S[:, -1]).all()
    recs = algo.recommend_batch(queries, 20)
    assert len(recs) == len(queries) and len(recs[-1]) == 0
    for b in range(len(queries) - 1):
        ref = oracle.argtopn(S[:, b], 20)
        assert np.array_equal(recs[b].numbers(vocabulary=algo.items), ref)  # selection is exact on the device scores
        assert np.array_equal(recs[b].ids(), ml_small.item_ids[ref])
        assert np.array_equal(recs[b].scores(), S[ref, b])

    # explicit-feedback model: biases enter the ranking (finalize_scores)
    bm = BiasedMFScorer(features=16, epochs=3)
    bm.train(ds, TrainingOptions(rng=7))
    qb = [RecQuery(user_id=ml_small.user_ids[u]) for u in (1, 2, 600)] + [
        RecQuery(user_id=None, query_items=ds.user_history(5))
    ]
    Sb = bm.score_matrix(qb).cpu().numpy()
    for b, q in enumerate(qb):
        host = bm(q, all_items).scores()
        assert np.allclose(Sb[:, b], host, rtol=2e-4, atol=2e-4), b
    rb = bm.recommend_batch(qb, 10)
    for b in range(len(qb)):
        assert np.array_equal(rb[b].numbers(vocabulary=bm.items), oracle.argtopn(Sb[:, b], 10))
