"""
The roofline arithmetic of bench.py against the figures of SURVEY.md §8d (algorithmic bytes per
half-epoch / per build at ML-25M shape), and the host-side helpers that do not need a GPU.
"""

import importlib.util
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("lk_bench", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lk_bench"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_als_algorithmic_bytes_match_survey():
    b = _bench()
    U, I, N, k = 162_541, 59_047, 25_000_095, 64
    # N(k s + 8) + 4(R+1) + R k 4 2 + k^2 4 per half-epoch (OtOr pass counted separately)
    fp32 = b.als_half_bytes(U, I, N, k, 4, False) + b.als_half_bytes(I, U, N, k, 4, False)
    bf16 = b.als_half_bytes(U, I, N, k, 2, False) + b.als_half_bytes(I, U, N, k, 2, False)
    assert abs(fp32 / 1e9 - 13.3) < 0.1  # "≈ 13.4 GB/epoch" with the OtOr pass
    assert abs(bf16 / 1e9 - 6.9) < 0.1  # "bf16 gather ≈ 7.0 GB"
    with_otor = b.als_half_bytes(U, I, N, k, 4, True) + b.als_half_bytes(I, U, N, k, 4, True)
    assert abs(with_otor / 1e9 - 13.4) < 0.1
    # 2.0 ms / 1.05 ms at the measured 6,571 GB/s
    assert abs(fp32 / 6571.2e9 * 1e3 - 2.03) < 0.05 and abs(bf16 / 6571.2e9 * 1e3 - 1.05) < 0.05


def test_knn_algorithmic_bytes_match_survey():
    b = _bench()
    # P = 1.278e10 products => 102 GB streamed, 15.6 ms at peak
    by = b.knn_build_bytes(12_764_118_364, 25_000_095, 162_541, 59_047, 1_180_940)
    assert abs(by / 1e9 - 102.3) < 0.5
    assert abs(by / 6571.2e9 * 1e3 - 15.6) < 0.2


def test_host_threads_ignores_omp_env(monkeypatch):
    b = _bench()
    monkeypatch.setenv("OMP_NUM_THREADS", "1")  # what torchrun exports to every rank
    import os

    assert b.host_threads() == len(os.sched_getaffinity(0)) >= 1


def test_committed_bench_lines_have_the_contract_keys():
    """The bench lines kept under profiles/ carry every key the driver's contract names."""
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"}  # fmt: skip
    for name in ("r01_final_bench_n1.json", "r01_final_bench_n2.json", "r01_final_bench_n4.json"):
        line = json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])
        assert need <= set(line), (name, need - set(line))
        assert line["vs_baseline"] is None and line["higher_is_better"] is False
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
        assert "workload" in line["config"]
    ref = json.loads((ROOT / "profiles" / "r01_final_bench_reference_arm.json").read_text().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0
    assert {"kind", "cores", "sample", "value"} <= set(ref["cpu_baseline"])


def test_solve_cholesky_property():
    """The reference's property test of `solve_cholesky` (tests/utils/test_math_solve.py:21-50) against the
    host fold-in solver that mirrors it (`lkpy_b200.als._solve_cholesky`)."""
    import numpy as np
    from pytest import approx

    from lkpy_b200.als import _solve_cholesky

    for seed in range(40):
        rng = np.random.RandomState(seed)
        size = int(rng.randint(2, 101))
        A = rng.randn(size, size) * 10
        b = rng.randn(size) * 10
        A = A * A
        xexp = np.linalg.lstsq(A, b, rcond=None)[0]
        F, y = A.T @ A, A.T @ b
        x = _solve_cholesky(F, y)
        assert x.shape == y.shape
        assert x == approx(xexp, rel=1.0e-3)
        assert F @ x == approx(y, rel=2.0e-6, abs=5.0e-9)
