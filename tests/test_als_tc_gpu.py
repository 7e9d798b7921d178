"""
GPU parity of the tensor-core ALS kernel (als_tc.cu: k = 64, bf16 gather,
unweighted / uniformly weighted Gram) against the oracle that rounds the gathered
rows to bf16 the same way, and against the SIMT kernel — for every solve variant
the library can be switched to.
"""

import numpy as np
import pytest
import torch

import oracle
from lkpy_b200 import _lib, data, engine

from helpers import rel_fro, small_synth

pytestmark = pytest.mark.gpu

# solve variants of the tensor-core kernel (switches of lk_set_option, read by launch_als_tc / dispatch)
VARIANTS = {
    "tc-gauss-jordan": {},  # default: block Gauss-Jordan in TMEM with tcgen05 trailing updates (chol_tc.cuh)
    "tc-cholesky": {"LK_ALS_GJ": 0},  # blocked Cholesky + block back substitution variant of the same
    "smem-solve": {"LK_ALS_TCS": 0},  # systems drained to shared memory, one warp per solve
    "smem-solve-wide": {"LK_ALS_TCS": 0, "LK_ALS_TC_INTERLEAVE": 0},  # one accumulator per 64 columns
}
DEFAULTS = {"LK_ALS_TC": 1, "LK_ALS_TCS": 1, "LK_ALS_GJ": 1, "LK_ALS_TC_INTERLEAVE": 1}


def _set(lk_options, variant):
    for key, val in {**DEFAULTS, **VARIANTS[variant]}.items():
        lk_options(key, val)


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


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("mode", ["implicit", "explicit"])
def test_tc_kernel_parity(cuda_lib, lk_options, mode, variant):
    _set(lk_options, variant)
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
        assert rel_fro(got, ref) < 1e-4, rel_fro(got, ref)  # north_star tolerance
        assert delta == pytest.approx(dref, rel=1e-3)
        empty = np.diff(csr.indptr) == 0
        assert np.all(got[empty] == 0.0)
        # the SIMT kernel on the same inputs agrees to rounding
        lk_options("LK_ALS_TC", 0)
        simt, _, _ = _run(mode, csr, this, other, 0.1)
        _set(lk_options, variant)
        assert rel_fro(got, simt) < 2e-5


@pytest.mark.parametrize("variant", ["tc-cholesky", "tc-gauss-jordan", "smem-solve"])
@pytest.mark.parametrize("mode", ["implicit", "explicit"])
def test_tc_split_rows_deterministic(cuda_lib, lk_options, variant, mode):
    _set(lk_options, variant)
    inter = small_synth(300, 200, 20000, seed=5)
    rng = np.random.default_rng(5)
    if mode == "implicit":
        _ui, iu = data.als_implicit_matrices(inter, 40.0)
    else:
        coo = inter.coo(rng.standard_normal(inter.nnz).astype(np.float32))
        iu = data.InteractionCSR.from_scipy(coo.T)
    p = (rng.standard_normal((300, 64)) * 0.1).astype(np.float32)
    q = (rng.standard_normal((200, 64)) * 0.1).astype(np.float32)
    a, _, plan = _run(mode, iu, q, p, 0.1, chunk_nnz=32)
    assert plan.n_split_rows > 0
    b, _, _ = _run(mode, iu, q, p, 0.1, chunk_nnz=32)
    assert np.array_equal(a.view(np.int32), b.view(np.int32))  # bit-reproducible
    ref, _ = _oracle(mode, iu, q, p, 0.1)
    assert rel_fro(a, ref) < 1e-4
    c, _, _ = _run(mode, iu, q, p, 0.1, chunk_nnz=1 << 20)
    assert rel_fro(a, c) < 1e-5


@pytest.mark.parametrize("variant", ["tc-cholesky", "tc-gauss-jordan"])
def test_tc_solve_badly_conditioned(cuda_lib, lk_options, variant):
    """Large Gram, small ridge (cond ~1e4): the tensor-core solve stays within a small factor of what
    f32 arithmetic can deliver (the f32 oracle's own distance to the f64 oracle)."""
    _set(lk_options, variant)
    inter = small_synth(400, 300, 60000, seed=9)
    ui, _ = data.als_implicit_matrices(inter, 40.0)
    rng = np.random.default_rng(9)
    p = (rng.standard_normal((400, 64)) * 0.1).astype(np.float32)
    q = rng.standard_normal((300, 64)).astype(np.float32)  # trained-scale-and-beyond item factors
    reg = 0.01
    got, _, _ = _run("implicit", ui, p, q, reg)
    ref, _ = _oracle("implicit", ui, p, q, reg)
    o32, _ = oracle.otor(oracle.bf16_round(q), reg)
    cpu32, _ = oracle.als_half("implicit", ui, p, q, otor_mat=o32, bf16_other=True)
    e_gpu, e_cpu = rel_fro(got, ref), rel_fro(cpu32, ref)
    assert e_gpu < max(1e-4, 3.0 * e_cpu), (e_gpu, e_cpu)


def test_tc_zero_weight_falls_back(cuda_lib, lk_options):
    """weight = 0 makes every confidence 0: (A / v) is undefined, the shared-memory solve takes over."""
    _set(lk_options, "tc-cholesky")
    inter = small_synth(200, 150, 5000, seed=3)
    ui, _ = data.als_implicit_matrices(inter, 0.0)
    rng = np.random.default_rng(3)
    p = (rng.standard_normal((200, 64)) * 0.1).astype(np.float32)
    q = (rng.standard_normal((150, 64)) * 0.1).astype(np.float32)
    got, _, _ = _run("implicit", ui, p, q, 0.1)
    ref, _ = _oracle("implicit", ui, p, q, 0.1)
    assert np.isfinite(got).all()
    assert rel_fro(got, ref) < 1e-4


def test_tc_non_uniform_weights_leave_the_bf16_fast_path(cuda_lib, lk_options):
    """use_ratings=True confidences are not uniform: als_tc.cu declines, als_tcx.cu (or, switched off, the
    SIMT kernel) takes the launch."""
    _set(lk_options, "tc-cholesky")
    inter = small_synth(400, 300, 15000, seed=8)
    ui, _ = data.als_implicit_matrices(inter, 40.0, use_ratings=True)
    rng = np.random.default_rng(8)
    p = (rng.standard_normal((400, 64)) * 0.1).astype(np.float32)
    q = (rng.standard_normal((300, 64)) * 0.1).astype(np.float32)
    got, _, plan = _run("implicit", ui, p, q, 0.1)
    assert not plan.vals_uniform
    ref, _ = _oracle("implicit", ui, p, q, 0.1)
    assert rel_fro(got, ref) < 1e-4


@pytest.mark.parametrize("variant", ["tc-cholesky", "tc-gauss-jordan", "smem-solve"])
def test_tc_not_positive_definite_is_reported(cuda_lib, lk_options, variant):
    """A system that is not positive definite (the reference's `ALS solve error`, implicit.rs:79) sets the
    status word and leaves the row untouched — also on the tensor-core path."""
    _set(lk_options, variant)
    inter = small_synth(60, 40, 600, seed=2)
    ui, _ = data.als_implicit_matrices(inter, 40.0)
    dev = _lib.require_device()
    k = 64
    dm = engine.DeviceCSR.from_host(ui, dev)
    plan = engine.ALSHalfPlan.create(dm, k)
    assert plan.vals_uniform
    this = torch.full((60, k), 0.5, device=dev)
    other = torch.zeros((40, k), device=dev, dtype=torch.bfloat16)
    otor = -torch.eye(k, device=dev)
    engine.als_half_epoch(plan, _lib.LK_ALS_IMPLICIT, this, other, otor=otor)
    torch.cuda.synchronize()
    assert int(plan.status.item()) > 0
    nonempty = torch.from_numpy(np.diff(ui.indptr) > 0).to(dev)
    assert torch.all(this[nonempty] == 0.5)  # failed solves do not write
