"""Shared helpers for the parity tests."""

from __future__ import annotations

import numpy as np

from lkpy_b200 import data


def implicit_init(rng: np.random.Generator, n_items: int, n_users: int, k: int):
    """Reference init: items first, then users; (0.01*N(0,1))^2 (als/_implicit.py:152-155)."""

    def one(n):
        m = rng.standard_normal((n, k), dtype=np.float32) * 0.01
        m *= m
        return m

    q = one(n_items)
    p = one(n_users)
    return p, q


def explicit_init(rng: np.random.Generator, n_items: int, n_users: int, k: int):
    """Reference init: unit-norm random rows (als/_explicit.py:105-108)."""

    def one(n):
        m = rng.standard_normal((n, k), dtype=np.float32)
        m /= np.linalg.norm(m, axis=1).reshape((n, 1))
        return m

    q = one(n_items)
    p = one(n_users)
    return p, q


def rel_fro(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def small_synth(n_users=600, n_items=400, nnz=20000, seed=7, ratings=True):
    return data.synth_interactions(n_users, n_items, nnz, seed, ratings=ratings)
