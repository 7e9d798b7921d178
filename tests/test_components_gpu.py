"""
GPU tests of the component surface (``lkpy_b200.als`` / ``lkpy_b200.knn``), modelled on the
reference's conformance suite ``lenskit.testing.ScorerTests``
(``src/lenskit/testing/_components.py:110-378``) and on ``tests/models/test_als_implicit.py``,
``test_als_explicit.py``, ``test_knn_item_item.py``: train on ml-latest-small, score known /
unknown users and items, empty inputs, pickle round trip, fold-in consistency, golden predictions.
"""

import pickle
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

import oracle
from lkpy_b200 import data
from lkpy_b200.als import BiasedMFScorer, ImplicitMFScorer
from lkpy_b200.components import Dataset, ItemList, RecQuery, TrainingOptions
from lkpy_b200.knn import ItemKNNScorer

from helpers import explicit_init, implicit_init, rel_fro

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def ml_ds(ml_small):
    return Dataset(ml_small)


def _oracle_implicit(inter, k, epochs, seed, weight=40.0, reg=0.1):
    """The reference training loop (als/_common.py:209-256) run on the CPU oracle."""
    ui, iu = data.als_implicit_matrices(inter, weight)
    p, q = implicit_init(np.random.default_rng(seed), inter.n_items, inter.n_users, k)
    for _ in range(epochs):
        o32, _ = oracle.otor(q, reg)
        p, _ = oracle.als_half("implicit", ui, p, q, otor_mat=o32)
        o32, _ = oracle.otor(p, reg)
        q, _ = oracle.als_half("implicit", iu, q, p, otor_mat=o32)
    return p, q


def test_implicit_mf_train_and_score(cuda_lib, ml_ds, ml_small):
    algo = ImplicitMFScorer(features=32, epochs=4)
    assert not algo.is_trained()
    algo.train(ml_ds, TrainingOptions(rng=42))
    assert algo.is_trained()
    assert algo.user_embeddings.shape == (ml_small.n_users, 32)
    assert algo.item_embeddings.shape == (ml_small.n_items, 32)
    assert algo.user_embeddings.dtype == np.float32
    assert len(algo.training_log) == 4 and all(m["deltaP"] > 0 for m in algo.training_log)

    # same init, same loop on the CPU oracle: the models must rank alike.  (Element-wise factor
    # parity over several epochs is not meaningful — SURVEY.md §7: f32 vs f64 drifts to 1e-3.)
    p, q = _oracle_implicit(ml_small, 32, 4, 42)
    rng = np.random.default_rng(0)
    for u in rng.choice(ml_small.n_users, 25, replace=False):
        s_gpu = algo.item_embeddings @ algo.user_embeddings[u]
        s_ref = q @ p[u]
        assert np.corrcoef(s_gpu, s_ref)[0, 1] > 0.999
        assert rel_fro(s_gpu, s_ref) < 2e-2

    # known user, known + unknown items
    uid = ml_small.user_ids[10]
    items = ItemList([ml_small.item_ids[5], ml_small.item_ids[77], -9999])
    res = algo(RecQuery(user_id=uid), items)
    sc = res.scores()
    assert np.isfinite(sc[:2]).all() and np.isnan(sc[2])
    assert sc[0] == pytest.approx(float(algo.item_embeddings[5] @ algo.user_embeddings[10]), rel=1e-5)
    # unknown user without history -> all NaN (als/_common.py:155-158)
    assert np.isnan(algo(RecQuery(user_id=-5), items).scores()).all()
    # empty item list
    assert len(algo(RecQuery(user_id=uid), ItemList([]))) == 0

    # fold-in from the user's history agrees with the trained row in ranking
    # (test_als_implicit.py:221-274 asks for Kendall tau >= 0.5)
    hist = ml_ds.user_history(10)
    res_fold = algo(RecQuery(user_id=None, query_items=hist), ItemList(ml_small.item_ids[:500]))
    res_train = algo(RecQuery(user_id=uid), ItemList(ml_small.item_ids[:500]))
    assert np.corrcoef(res_fold.scores(), res_train.scores())[0, 1] > 0.9

    # pickle round trip (ScorerTests: approx(abs=1e-3))
    clone = pickle.loads(pickle.dumps(algo))
    assert np.array_equal(clone.item_embeddings, algo.item_embeddings)
    assert np.allclose(clone(RecQuery(user_id=uid), items).scores()[:2], sc[:2], atol=1e-3)


def test_implicit_mf_options(cuda_lib, ml_ds, ml_small):
    a = ImplicitMFScorer(features=20, epochs=2, user_embeddings=False, regularization=(0.05, 0.2))
    a.train(ml_ds, TrainingOptions(rng=1))
    assert a.user_embeddings is None and a.users is None
    assert a.item_embeddings.shape == (ml_small.n_items, 20)
    # use_ratings=True makes the confidences non-uniform (SIMT kernel); bf16 gather runs too
    b = ImplicitMFScorer(features=64, epochs=2, use_ratings=True, gather_dtype="bfloat16")
    b.train(ml_ds, TrainingOptions(rng=1))
    assert np.isfinite(b.item_embeddings).all()
    # retrain=False keeps a trained model
    before = b.item_embeddings.copy()
    b.train(ml_ds, TrainingOptions(rng=2, retrain=False))
    assert np.array_equal(before, b.item_embeddings)


def test_biased_mf_train_and_score(cuda_lib, ml_ds, ml_small):
    algo = BiasedMFScorer(features=32, epochs=5)
    algo.train(ml_ds, TrainingOptions(rng=42))
    assert algo.bias.global_bias == pytest.approx(ml_small.ratings.mean(), rel=1e-6)

    # the same loop on the CPU oracle (bias model identical by construction)
    resid = algo.bias.transform(ml_small)
    coo = ml_small.coo(resid)
    ui, iu = data.InteractionCSR.from_scipy(coo), data.InteractionCSR.from_scipy(coo.T)
    p, q = explicit_init(np.random.default_rng(42), ml_small.n_items, ml_small.n_users, 32)
    for _ in range(5):
        p, _ = oracle.als_half("explicit", ui, p, q, reg=0.1)
        q, _ = oracle.als_half("explicit", iu, q, p, reg=0.1)
    pred_gpu = np.einsum("ij,ij->i", algo.user_embeddings[ml_small.users], algo.item_embeddings[ml_small.items])
    pred_ref = np.einsum("ij,ij->i", p[ml_small.users], q[ml_small.items])
    assert rel_fro(pred_gpu, pred_ref) < 1e-2
    rmse = np.sqrt(np.mean((pred_gpu - resid) ** 2))
    assert rmse < 0.8  # training-set residual RMSE of a 32-factor model

    uid = ml_small.user_ids[3]
    items = ItemList(ml_small.item_ids[:50])
    sc = algo(RecQuery(user_id=uid), items).scores()
    assert np.isfinite(sc).all() and 0.0 < np.mean(sc) < 6.0
    # fold-in path with ratings
    sc2 = algo(RecQuery(user_id=None, query_items=ml_ds.user_history(3)), items).scores()
    assert np.corrcoef(sc, sc2)[0, 1] > 0.8


@pytest.mark.parametrize("feedback", ["explicit", "implicit"])
@pytest.mark.parametrize("save_nbrs", [None, 50])
def test_item_knn_train(cuda_lib, ml_ds, ml_small, feedback, save_nbrs):
    algo = ItemKNNScorer(k=20, save_nbrs=save_nbrs, feedback=feedback)
    algo.train(ml_ds)
    assert algo.is_trained()
    ui, iu, means = data.knn_item_matrices(ml_small, feedback == "explicit")
    ref = oracle.knn_build(ui, iu, 1e-6, save_nbrs)
    sm = algo.sim_matrix
    assert sm.indptr.dtype == np.int64  # LargeList offsets (item_score.rs:113-118)
    assert np.array_equal(sm.indptr, ref.indptr)
    assert np.array_equal(sm.indices, ref.indices)
    assert np.array_equal(sm.values.view(np.int32), ref.data.view(np.int32))
    assert np.array_equal(algo.item_counts, np.diff(ref.indptr))
    if feedback == "explicit":
        assert np.array_equal(algo.item_means, means)
        assert np.all(sm.values > 0) and np.all(sm.values < 1 + 1e-6)  # test_knn_item_item.py:135-137
    else:
        assert algo.item_means is None


def test_item_knn_known_preds(cuda_lib, ml_ds, ml_small):
    """test_ii_known_preds (test_knn_item_item.py:413-453) through the component API."""
    algo = ItemKNNScorer(k=20, min_sim=1.0e-6)
    algo.train(ml_ds)
    known = pd.read_csv(GOLD / "item-item-preds.csv")
    uidx = {u: i for i, u in enumerate(ml_small.user_ids)}
    queries, targets, expected = [], [], []
    for uid, grp in known.groupby("user_id"):
        queries.append(ml_ds.user_history(uidx[uid]))
        targets.append(ItemList(grp.item_id.values))
        expected.append(grp.prediction.values)
    outs = algo.score_batch(queries, targets)
    err = np.abs(np.concatenate([o.scores() for o in outs]) - np.concatenate(expected))
    assert len(err) == 1288 and not np.isnan(err).any()
    assert (err > 1e-5).sum() <= 3
    # the per-query call gives the same numbers and carries nbr_counts
    one = algo(RecQuery(query_items=queries[0]), targets[0])
    assert np.array_equal(one.scores(), outs[0].scores())
    assert one.field("nbr_counts").max() <= 20
    # no history -> NaN scores (knn/item.py:238-245)
    assert np.isnan(algo(RecQuery(user_id=1), targets[0]).scores()).all()


def test_item_knn_implicit_topk_sum(cuda_lib, ml_ds, ml_small):
    """test_ii_implicit_large (test_knn_item_item.py:373-410): score == sum of the k largest sims."""
    nbrs = 5
    algo = ItemKNNScorer(k=nbrs, feedback="implicit")
    algo.train(ml_ds)
    mat = algo.sim_matrix.to_scipy().toarray()
    rng = np.random.default_rng(3)
    for u in rng.choice(ml_small.n_users, 10, replace=False):
        hist = ml_ds.user_history(int(u))
        tgt = ItemList(rng.choice(ml_small.item_ids, 200, replace=False))
        res = algo(RecQuery(query_items=hist), tgt)
        rows = mat[hist.numbers(vocabulary=algo.items), :]
        for iid, score in zip(tgt.ids(), res.scores()):
            col = rows[:, algo.items.number(iid)]
            top = np.sort(col[col > 0])[::-1][:nbrs]
            if len(top) == 0:
                assert np.isnan(score)
            else:
                assert score == pytest.approx(float(top.sum()), rel=1e-5)


def test_item_knn_pickle(cuda_lib, ml_ds, ml_small):
    algo = ItemKNNScorer(k=20, save_nbrs=30)
    algo.train(ml_ds)
    clone = pickle.loads(pickle.dumps(algo))
    assert np.array_equal(clone.sim_matrix.values, algo.sim_matrix.values)
    assert np.array_equal(clone.item_means, algo.item_means)
    hist = ml_ds.user_history(5)
    tgt = ItemList(ml_small.item_ids[:300])
    a = algo(RecQuery(query_items=hist), tgt).scores()
    b = clone(RecQuery(query_items=hist), tgt).scores()
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
