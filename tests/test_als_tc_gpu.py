"""
GPU parity of the tensor-core ALS kernel (als_tc.cu: k = 64, bf16 gather,
unweighted / uniformly weighted Gram) against the oracle that rounds the gathered
rows to bf16 the same way, and against the SIMT kernel.
"""

import numpy as np
import pytest
import torch

import oracle
from lkpy_b200 import _lib, data, engine

from helpers import rel_fro, small_synth

pytestmark = pytest.mark.gpu


def _run(mode, csr, this, other, reg, chunk_nnz=engine.DEFAULT_CHUNK_NNZ):
    dev = _lib.require_device()
    dm = engine.DeviceCSR.from_host(csr, dev)
    k = this.shape[1]
    plan = engine.ALSHalfPlan.create(dm, k, chunk_nnz)
    d_this = torch.from_numpy(this.copy()).to(dev)
    d_other = torch.from_numpy(other).to(dev)
    obf = torch.empty_like(d_other, dtype=torch.bfloat16)
    ws = engine.OtorWorkspace.create(k, dev)
    otor = engine.als_otor(d_other, reg, ws, obf)  # also fills the bf16 copy
    engine.als_half_epoch(
        plan, _lib.LK_ALS_IMPLICIT if mode == "implicit" else _lib.LK_ALS_EXPLICIT, d_this, obf,
        otor=otor if mode == "implicit" else None, reg=reg,
    )  # fmt: skip
    torch.cuda.synchronize()
    assert int(plan.status.item()) == 0
    return d_this.cpu().numpy(), float(np.sqrt(plan.sqdelta.item())), plan


def _oracle(mode, csr, this, other, reg):
    _o32, o64 = oracle.otor(oracle.bf16_round(other), reg)
    return oracle.als_half_f64(mode, csr, this, other, otor_mat=o64, reg=reg, bf16_other=True)


# generation "2" = als_tcr.cu (register-resident solve, the default), "1" = als_tc.cu
@pytest.mark.parametrize("gen,interleave", [("2", "1"), ("1", "1"), ("1", "0")])
@pytest.mark.parametrize("mode", ["implicit", "explicit"])
def test_tc_kernel_parity(cuda_lib, monkeypatch, mode, gen, interleave):
    monkeypatch.setenv("LK_ALS_TC", gen)
    monkeypatch.setenv("LK_ALS_TC_INTERLEAVE", interleave)
    inter = small_synth(900, 500, 40000, seed=21)
    rng = np.random.default_rng(21)
    k = 64
    p = (rng.standard_normal((inter.n_users, k)) * 0.1).astype(np.float32)
    q = (rng.standard_normal((inter.n_items, k)) * 0.1).astype(np.float32)
    if mode == "implicit":
        ui, iu = data.als_implicit_matrices(inter, 40.0)
    else:
        coo = inter.coo(rng.standard_normal(inter.nnz).astype(np.float32))
        ui, iu = data.InteractionCSR.from_scipy(coo), data.InteractionCSR.from_scipy(coo.T)
    for csr, this, other in ((ui, p, q), (iu, q, p)):
        got, delta, plan = _run(mode, csr, this, other, 0.1)
        assert plan.vals_uniform == (mode == "implicit")
        ref, dref = _oracle(mode, csr, this, other, 0.1)
        assert rel_fro(got, ref) < 1e-4, rel_fro(got, ref)
        assert delta == pytest.approx(dref, rel=1e-3)
        empty = np.diff(csr.indptr) == 0
        assert np.all(got[empty] == 0.0)
        # the SIMT kernel on the same inputs agrees to rounding
        monkeypatch.setenv("LK_ALS_TC", "0")
        simt, _, _ = _run(mode, csr, this, other, 0.1)
        monkeypatch.setenv("LK_ALS_TC", gen)
        assert rel_fro(got, simt) < 2e-5


@pytest.mark.parametrize("gen", ["2", "1"])
def test_tc_split_rows_deterministic(cuda_lib, monkeypatch, gen):
    monkeypatch.setenv("LK_ALS_TC", gen)
    inter = small_synth(300, 200, 20000, seed=5)
    _ui, iu = data.als_implicit_matrices(inter, 40.0)
    rng = np.random.default_rng(5)
    p = (rng.standard_normal((300, 64)) * 0.1).astype(np.float32)
    q = (rng.standard_normal((200, 64)) * 0.1).astype(np.float32)
    a, _, plan = _run("implicit", iu, q, p, 0.1, chunk_nnz=32)
    assert plan.n_split_rows > 0
    b, _, _ = _run("implicit", iu, q, p, 0.1, chunk_nnz=32)
    assert np.array_equal(a.view(np.int32), b.view(np.int32))
    ref, _ = _oracle("implicit", iu, q, p, 0.1)
    assert rel_fro(a, ref) < 1e-4
    c, _, _ = _run("implicit", iu, q, p, 0.1, chunk_nnz=1 << 20)
    assert rel_fro(a, c) < 1e-5


def test_tc_non_uniform_weights_fall_back(cuda_lib, monkeypatch):
    """use_ratings=True confidences are not uniform: the SIMT kernel must take the launch."""
    monkeypatch.setenv("LK_ALS_TC", "2")
    inter = small_synth(400, 300, 15000, seed=8)
    ui, _ = data.als_implicit_matrices(inter, 40.0, use_ratings=True)
    rng = np.random.default_rng(8)
    p = (rng.standard_normal((400, 64)) * 0.1).astype(np.float32)
    q = (rng.standard_normal((300, 64)) * 0.1).astype(np.float32)
    got, _, plan = _run("implicit", ui, p, q, 0.1)
    assert not plan.vals_uniform
    ref, _ = _oracle("implicit", ui, p, q, 0.1)
    assert rel_fro(got, ref) < 1e-4
