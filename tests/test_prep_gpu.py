"""
Device-side input preparation: lk_knn_prep_columns + the CSR/CSC plumbing of
lkpy_b200.prep must hand the item-kNN build the very bits the reference's SciPy host prep produces
(knn/item.py:202-228), and the transposed CSR pair of the ALS trainers (als/_common.py:216-219).
"""

import numpy as np
import pytest
import torch

from lkpy_b200 import _lib, data, prep

from helpers import small_synth

pytestmark = pytest.mark.gpu


def _up(inter, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    return t(inter.users), t(inter.items), t(inter.ratings)


def _same(dcsr, hcsr):
    assert np.array_equal(dcsr.indptr.cpu().numpy(), hcsr.indptr)
    assert np.array_equal(dcsr.indices.cpu().numpy(), hcsr.indices)
    assert np.array_equal(dcsr.values.cpu().numpy().view(np.int32), hcsr.values.view(np.int32))


@pytest.mark.parametrize("explicit", [True, False])
def test_knn_prep_bit_identical_ml_small(cuda_lib, ml_small, explicit):
    dev = _lib.require_device()
    ui, iu, means = data.knn_item_matrices(ml_small, explicit)
    u, i, r = _up(ml_small, dev)
    d_ui, d_iu, d_means = prep.knn_item_matrices_device(u, i, r if explicit else None, ml_small.n_users, ml_small.n_items, explicit)
    _same(d_ui, ui)
    _same(d_iu, iu)
    if explicit:
        assert np.array_equal(d_means.cpu().numpy().view(np.int32), means.view(np.int32))
    else:
        assert d_means is None


def test_knn_prep_bit_identical_long_columns(cuda_lib):
    """Columns of every length class of NumPy's pairwise summation (< 8, <= 128, several split levels)
    and empty items."""
    dev = _lib.require_device()
    base = small_synth(30000, 300, 400000, seed=31)  # hot items have thousands of ratings
    # plus a tail of rarely rated items (1..9 ratings) and two items nobody rated
    rng = np.random.default_rng(31)
    tu, ti = [], []
    for extra, n in enumerate([1, 2, 3, 5, 7, 8, 9, 0, 0]):
        tu.append(np.sort(rng.choice(30000, n, replace=False)))
        ti.append(np.full(n, 300 + extra))
    users = np.concatenate([base.users] + tu).astype(np.int32)
    items = np.concatenate([base.items] + ti).astype(np.int32)
    ratings = np.concatenate([base.ratings, data.ML_RATING_VALUES[rng.integers(0, 10, len(users) - base.nnz)]])
    order = np.lexsort((items, users))
    inter = data.Interactions(users[order], items[order], ratings[order].astype(np.float32), 30000, 309)
    counts = np.bincount(inter.items, minlength=inter.n_items)
    assert counts.max() > 5000 and (counts < 8).any() and (counts == 0).any()
    for explicit in (True, False):
        ui, iu, means = data.knn_item_matrices(inter, explicit)
        u, i, r = _up(inter, dev)
        d_ui, d_iu, d_means = prep.knn_item_matrices_device(u, i, r if explicit else None, inter.n_users, inter.n_items, explicit)
        _same(d_ui, ui)
        _same(d_iu, iu)
        if explicit:
            assert np.array_equal(d_means.cpu().numpy().view(np.int32), means.view(np.int32))


def test_csr_pair_matches_scipy(cuda_lib):
    dev = _lib.require_device()
    inter = small_synth(900, 500, 40000, seed=8)
    ui, iu = data.als_implicit_matrices(inter, 40.0, use_ratings=True)
    u, i, r = _up(inter, dev)
    d_ui, d_iu, _perm = prep.coo_to_csr_pair(u, i, r * 40.0, inter.n_users, inter.n_items)
    _same(d_ui, ui)
    _same(d_iu, iu)
    assert np.array_equal(d_ui.h_indptr, ui.indptr) and np.array_equal(d_iu.h_indptr, iu.indptr)


def test_item_knn_train_device_prep_equals_host_prep(cuda_lib, ml_small):
    from lkpy_b200.components import Dataset
    from lkpy_b200.knn import ItemKNNScorer

    a = ItemKNNScorer(save_nbrs=20, prep="device")
    b = ItemKNNScorer(save_nbrs=20, prep="host")
    a.train(Dataset(ml_small))
    b.train(Dataset(ml_small))
    assert np.array_equal(a.sim_matrix.indptr, b.sim_matrix.indptr)
    assert np.array_equal(a.sim_matrix.indices, b.sim_matrix.indices)
    assert np.array_equal(a.sim_matrix.values.view(np.int32), b.sim_matrix.values.view(np.int32))
    assert np.array_equal(a.item_means.view(np.int32), b.item_means.view(np.int32))


def test_device_generator_shape(cuda_lib):
    """The device generator has the NumPy generator's construction: sorted unique pairs, power-law items,
    heavy-tailed users, ratings from the ML pmf."""
    dev = _lib.require_device()
    u, i, r = prep.synth_interactions_device(20000, 6000, 1_000_000, seed=5, device=dev)
    assert u.numel() == i.numel() == r.numel() == 1_000_000
    keys = u.long() * 6000 + i.long()
    assert bool(torch.all(keys[1:] > keys[:-1]))  # sorted by (user, item), no duplicates
    ref = data.synth_interactions(20000, 6000, 1_000_000, seed=5)
    ic = torch.bincount(i.long(), minlength=6000).cpu().numpy()
    rc = np.bincount(ref.items, minlength=6000)
    assert 0.8 < np.sort(ic)[-10:].sum() / np.sort(rc)[-10:].sum() < 1.25  # same head of the item distribution
    assert 0.8 < np.median(ic) / np.median(rc) < 1.25
    uc = torch.bincount(u.long(), minlength=20000).cpu().numpy()
    assert 0.8 < np.median(uc) / np.median(np.bincount(ref.users, minlength=20000)) < 1.25
    assert set(np.unique(r.cpu().numpy())) <= set(data.ML_RATING_VALUES.tolist())
