#!/usr/bin/env python
"""
bench.py — ALS-implicit epoch time and item-kNN build throughput on
ML-25M-shaped synthetic interactions (BASELINE.json metric / configs[1], [2]).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU restatement of the reference path

One JSON line on stdout (rank 0).  A "step" is one ALS epoch (user half-epoch +
item half-epoch, each = OtOr kernel + row-solve kernel, + the factor all-gather
when N > 1) over the resident CSR matrices.  `value` is the epoch time with
everything resident in HBM; `e2e` is the same epoch through the public trainer
API with the factor tables coming from / going back to pinned host memory inside
the timed region.  Riding along: "als_fp32" (the same epoch with fp32 gathered rows),
"knn" (item-kNN build items/s with its kernel roofline and CPU figure, batched
scoring users/s) and, at N=1, "recommend" (batched scoring + top-100 users/s with
the selection kernel's roofline and the per-query CPU figure).  "cpu_baseline" and
`--impl reference` time the CPU restatement of the reference path (oracle/) on a
bounded sample at its best thread count.  Only the JSON line goes to stdout.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "als_implicit_epoch_time_ml25m_k64"
UNIT = "ms"
K = 64
WEIGHT = 40.0
REG = 0.1
KNN_SAVE = 20
KNN_MIN_SIM = 1e-6


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, gpu_index: int = 0):
        self.rows: list[list[str]] = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        assert self.proc and self.proc.stdout
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout: float = 8.0) -> None:
        """nvidia-smi takes a few hundred ms to print its first row: block until it is sampling."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin: float | None = None, t_end: float | None = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = list(self.rows)
        window = "timed region"
        if t_begin is not None and t_end is not None:
            inside = [r for t, r in rows if t_begin <= t <= t_end + 0.02]
            if len(inside) >= 2:
                sel = inside
            else:  # region shorter than two sampling periods: use the samples under the same load around it
                sel = [r for t, r in rows if t_begin - 0.5 <= t <= t_end + 0.1]
                window = "timed region +/- warm-up epochs (region shorter than two 20 ms samples)"
        else:
            sel = [r for _, r in rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in sel:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "window": window,
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------
# algorithmic bytes
# ---------------------------------------------------------------------------


def als_half_bytes(n_rows: int, n_other: int, nnz: int, k: int, s: int, with_otor: bool) -> float:
    b = nnz * (k * s + 8) + 4 * (n_rows + 1) + n_rows * k * 4 * 2 + k * k * 4
    if with_otor:
        b += n_other * k * 4 + (n_other * k * 2 if s == 2 else 0)
    return float(b)


def knn_build_bytes(products: int, nnz: int, n_users: int, n_items: int, nnz_out: int) -> float:
    return float(8 * products + 8 * nnz + 4 * (n_items + n_users + 2) + 8 * nnz_out)


def load_peaks() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key: str, world: int) -> float | None:
    """DRAM bytes per launch from the committed ncu --set full capture (N=1, default workload only)."""
    f = ROOT / "profiles" / "r01_traffic.json"
    if world != 1 or not f.exists():
        return None
    return float(json.loads(f.read_text()).get(key, 0)) or None


# ---------------------------------------------------------------------------
# CPU baseline (oracle port of src/accel; the Rust crate cannot be built here)
# ---------------------------------------------------------------------------


def host_threads() -> int:
    """
    Host threads the CPU arm may use: the cores this process may run on.  (Not OMP_NUM_THREADS:
    torchrun exports OMP_NUM_THREADS=1 to every rank, which would silently turn the multi-threaded
    CPU baseline into a single-threaded one whenever the bench is launched under it.)
    """
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def _slice_rows(csr, lo: int, hi: int):
    from lkpy_b200.data import InteractionCSR

    a, b = int(csr.indptr[lo]), int(csr.indptr[hi])
    return InteractionCSR(
        (csr.indptr[lo : hi + 1] - a).astype(np.int64), csr.indices[a:b], csr.values[a:b], (hi - lo, csr.shape[1])
    )


def cpu_als_epoch_estimate(ui, iu, p, q, budget_s: float, threads: int) -> dict:
    """
    Time the oracle (BLAS sgemm Gram + LAPACK sposv) on leading row ranges of both halves and
    scale by nnz to a full epoch.  The thread count is the fastest of {all host threads the BLAS
    allows, 32, 16, 8 (the reference's default, schemas/settings.py:183)} on a calibration slice:
    on many-core hosts the row-parallel loop stops scaling well before all cores are busy.
    """
    import oracle

    oracle.use_scipy_blas(True)
    try:
        out = {}
        total = 0.0
        parts = []
        top = oracle.blas_threads(threads)
        cands = sorted({t for t in (top, 32, 16, 8, 4) if t <= top}, reverse=True)
        used = []
        for name, csr, this, other in (("user", ui, p, q), ("item", iu, q, p)):
            o32 = (other.T @ other + np.eye(K, dtype=np.float32) * REG).astype(np.float32)
            # calibrate on ~3% of the nonzeros (thread count + rate), then size the sample to the budget
            nnz = csr.nnz
            rows_cal = max(int(np.searchsorted(csr.indptr, nnz // 20)), 16)
            cal = _slice_rows(csr, 0, rows_cal)
            best_t, t_cal = cands[0], float("inf")
            for th in cands:
                for _rep in range(2):  # best of two: the first call at a new thread count pays the pool start-up
                    t0 = time.perf_counter()
                    oracle.als_half("implicit", cal, this[:rows_cal], other, otor_mat=o32, threads=th)
                    dt = time.perf_counter() - t0
                    if dt < t_cal:
                        best_t, t_cal = th, dt
            used.append(best_t)
            nnz_cal = int(csr.indptr[rows_cal])
            want_nnz = int(min(nnz, nnz_cal * (budget_s / 2) / max(t_cal, 1e-6)))
            rows = int(np.searchsorted(csr.indptr, want_nnz))
            rows = min(max(rows, rows_cal), csr.shape[0])
            t0 = time.perf_counter()
            oracle.als_half("implicit", _slice_rows(csr, 0, rows), this[:rows], other, otor_mat=o32, threads=best_t)
            t = time.perf_counter() - t0
            nnz_s = int(csr.indptr[rows])
            est = t * nnz / max(nnz_s, 1)
            parts.append(f"{name} half: {rows} rows / {nnz_s} nnz in {t:.2f}s on {best_t} threads")
            total += est
        out["epoch_ms"] = total * 1e3
        out["threads"] = max(used)
        out["sample"] = "; ".join(parts) + f"; scaled by nnz to the full epoch; thread counts tried {cands}"
        return out
    finally:
        oracle.use_scipy_blas(False)


def cpu_knn_estimate(kui, kiu, cost: np.ndarray, budget_s: float, threads: int) -> dict:
    import oracle

    n_items = kiu.shape[0]
    total_cost = float(cost.sum())
    rows_cal = max(64, n_items // 400)
    t0 = time.perf_counter()
    oracle.knn_build(kui, kiu, KNN_MIN_SIM, KNN_SAVE, rows=(0, rows_cal), threads=threads)
    t_cal = time.perf_counter() - t0
    c_cal = float(cost[:rows_cal].sum())
    want = c_cal * budget_s / max(t_cal, 1e-6)
    rows = int(np.searchsorted(np.cumsum(cost), want))
    rows = min(max(rows, rows_cal), n_items)
    t0 = time.perf_counter()
    oracle.knn_build(kui, kiu, KNN_MIN_SIM, KNN_SAVE, rows=(0, rows), threads=threads)
    t = time.perf_counter() - t0
    c = float(cost[:rows].sum())
    est = t * total_cost / max(c, 1.0)
    return {
        "build_items_per_s": n_items / est,
        "sample": f"items 0..{rows} ({c / total_cost:.2%} of the products) in {t:.2f}s, scaled by product count",
    }


# ---------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------


def make_data():
    from lkpy_b200 import data

    t0 = time.time()
    inter = data.synth_interactions(**data.ML25M_SHAPE)
    log(f"[bench] synthetic ML-25M-shaped data: {inter.n_users}x{inter.n_items}, nnz {inter.nnz} ({time.time() - t0:.1f}s)")
    return inter


def run_reference(args, rank: int) -> None:
    """--impl reference: the CPU path (oracle port; Rust unavailable) on a bounded sample."""
    if rank != 0:
        return
    import oracle
    from lkpy_b200 import data

    inter = make_data()
    ui, iu = data.als_implicit_matrices(inter, WEIGHT)
    rng = np.random.default_rng(0)
    p = (rng.standard_normal((inter.n_users, K)) * 0.1).astype(np.float32)
    q = (rng.standard_normal((inter.n_items, K)) * 0.1).astype(np.float32)
    threads = host_threads()
    per_step = max(4.0, 120.0 / max(args.steps + args.warmup, 1))
    vals = []
    sample = ""
    for s in range(args.warmup + args.steps):
        r = cpu_als_epoch_estimate(ui, iu, p, q, per_step, threads)
        threads = r["threads"]
        if s >= args.warmup:
            vals.append(r["epoch_ms"])
        sample = r["sample"]
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ML-25M-shaped synthetic ImplicitMFScorer features=64 (configs[1])",
                   "n_users": inter.n_users, "n_items": inter.n_items, "nnz": inter.nnz},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    emit(line)


_JSON_OUT = None


def emit(line: dict) -> None:
    """The one JSON line goes to the real stdout; everything else (NCCL's version banner, library
    chatter written to fd 1) was redirected to stderr by main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget per path")
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--score-users", type=int, default=2048)
    ap.add_argument("--variants", default="bf16,fp32", help="ALS gather dtypes to time (profiling runs pass one)")
    ap.add_argument("--profile", action="store_true", help="under ncu: honour a small --warmup, skip e2e")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch

    from lkpy_b200 import _build, _lib, data, engine
    from lkpy_b200.als import ImplicitMFScorer, ImplicitMFTrainer
    from lkpy_b200.components import Dataset, TrainingOptions

    _build.build()
    _lib.lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = load_peaks()

    inter = make_data()
    ds = Dataset(inter)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    results: dict = {}
    launches = 0
    variants = [v for v in args.variants.split(",") if v]
    for tag, gdt in (("bf16", "bfloat16"), ("fp32", "float32")):
        if tag not in variants:
            continue
        if world > 1:
            from lkpy_b200.parallel import ShardedImplicitMFTrainer

            scorer = ImplicitMFScorer(features=K, epochs=1, regularization=REG, weight=WEIGHT, gather_dtype=gdt)
            tr = ShardedImplicitMFTrainer(scorer, ds, TrainingOptions(rng=42))
        else:
            scorer = ImplicitMFScorer(features=K, epochs=1, regularization=REG, weight=WEIGHT, gather_dtype=gdt)
            tr = ImplicitMFTrainer(scorer, ds, TrainingOptions(rng=42))
        # one epoch from the reference init first, so timed epochs see trained-scale factors
        for _ in range(args.warmup if args.profile else max(args.warmup, 3)):
            tr.train_epoch_device()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and tag == "bf16":
            sampler.start()
            sampler.wait_first()
        if tag == "bf16":
            for _ in range(3):  # every rank: the GPU stays under the same load while the sampler spins up
                tr.train_epoch_device()
        tr.kernel_events = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_begin = time.time()
        e0.record()
        for _ in range(args.steps):
            tr.train_epoch_device()
        e1.record()
        barrier()
        t_end = time.time()
        clocks = sampler.stop(t_begin, t_end) if (rank == 0 and tag == "bf16") else None
        ms = e0.elapsed_time(e1) / args.steps
        kern_ms = [a.elapsed_time(b) for a, b in tr.kernel_events]
        tr.kernel_events = None
        tr._raise_on_status()
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        s = 2 if tag == "bf16" else 4
        nnz = inter.nnz
        # per GPU: each rank gathers for its row shard only (1/world of the nonzeros and rows)
        alg = (
            als_half_bytes(inter.n_users, inter.n_items, nnz, K, s, False)
            + als_half_bytes(inter.n_items, inter.n_users, nnz, K, s, False)
        ) / world
        per_epoch_kernel_ms = float(np.sum(kern_ms)) / args.steps if kern_ms else float("nan")
        achieved = alg / (per_epoch_kernel_ms * 1e-3) / 1e9 if kern_ms else float("nan")
        results[tag] = {
            "ms_per_epoch": ms,
            "solve_kernel_ms_per_epoch": per_epoch_kernel_ms,
            "roofline": {
                "bound": "hbm",
                "kernel": ("als_tc_kernel" if tag == "bf16" else "als_half_kernel") + " (user + item launch of one epoch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic("als_tc_kernel_epoch_bytes", world) if tag == "bf16" else None,
                "algorithmic_bytes": alg, "peak_source": peak_src,
            },
            "clocks": clocks,
        }  # fmt: skip
        launches = 6 * args.steps
        log(f"[bench] ALS {tag}: {ms:.3f} ms/epoch (solve kernels {per_epoch_kernel_ms:.3f} ms), "
            f"{achieved:.0f} GB/s algorithmic = {achieved / peak:.3f} of {peak_src}")

        # ---- end to end through the trainer API: factors from / to pinned host memory
        if tag == "bf16" and not args.profile:
            hp = torch.from_numpy(scorer.user_embeddings).pin_memory()
            hq = torch.from_numpy(scorer.item_embeddings).pin_memory()
            for _ in range(2):
                tr.train_epoch_e2e(hp, hq)
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.steps):
                tr.train_epoch_e2e(hp, hq)
            e1.record()
            barrier()
            wall = (time.perf_counter() - t0) * 1e3 / args.steps
            e2e_ms = e0.elapsed_time(e1) / args.steps
            nbytes = (hp.numel() + hq.numel()) * 4
            h2d, d2h = nbytes, nbytes + 16
            api = "ImplicitMFTrainer.train_epoch_e2e: H2D factor tables, epoch, D2H factor tables + deltas; CSR resident"
            if world > 1:
                import torch.distributed as dist

                # every rank uploads the rows it owns, rank 0 reads the whole model back: sum over ranks
                t = torch.tensor([tr.e2e_bytes[0], tr.e2e_bytes[1], max(e2e_ms, wall)], device=dev, dtype=torch.float64)
                tmax = t.clone()
                dist.all_reduce(t)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                h2d, d2h = int(t[0].item()), int(t[1].item())
                e2e_ms = wall = float(tmax[2].item())
                api = ("ShardedImplicitMFTrainer.train_epoch_e2e: each rank uploads its row shards, NVLink all-gather, "
                       "epoch, rank 0 reads the whole model back (others their shards); bytes summed over ranks")
            results["e2e"] = {
                "value": max(e2e_ms, wall), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": api,
            }  # fmt: skip
            log(f"[bench] ALS e2e: {results['e2e']['value']:.3f} ms/epoch")
            if rank == 0 and world == 1 and not args.no_knn:
                results["recommend"] = bench_recommend(args, tr, dev, peak, peak_src)
        del tr, scorer
        torch.cuda.empty_cache()

    knn = None
    if not args.no_knn:
        knn = bench_knn(args, inter, dev, peak, peak_src, rank, world)

    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        import oracle

        threads = host_threads()
        ui, iu = data.als_implicit_matrices(inter, WEIGHT)
        rng = np.random.default_rng(0)
        p = (rng.standard_normal((inter.n_users, K)) * 0.1).astype(np.float32)
        q = (rng.standard_normal((inter.n_items, K)) * 0.1).astype(np.float32)
        r = cpu_als_epoch_estimate(ui, iu, p, q, args.cpu_seconds, threads)
        cpu = {"value": r["epoch_ms"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"]}
        log(f"[bench] CPU baseline ({r['threads']} threads): {r['epoch_ms']:.0f} ms/epoch  [{r['sample']}]")
        if knn is not None:
            kc = cpu_knn_estimate(knn.pop("_kui"), knn.pop("_kiu"), knn.pop("_cost"), args.cpu_seconds, threads)
            knn["cpu_baseline"] = {"value": kc["build_items_per_s"], "unit": "items/s", "cores": threads,
                                   "kind": "port", "sample": kc["sample"]}  # fmt: skip
            log(f"[bench] CPU kNN build: {kc['build_items_per_s']:.0f} items/s [{kc['sample']}]")
    if knn is not None:
        for k_ in ("_kui", "_kiu", "_cost"):
            knn.pop(k_, None)

    if rank == 0:
        head = results.get("bf16") or results["fp32"]
        line = {
            "metric": METRIC, "value": head["ms_per_epoch"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_epoch"],
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (bf16-stored gather operand)", "data": "synthetic",
            "config": {
                "workload": "ML-25M-shaped synthetic ImplicitMFScorer features=64 bf16 gather (BASELINE configs[1])",
                "n_users": inter.n_users, "n_items": inter.n_items, "nnz": inter.nnz, "features": K,
                "weight": WEIGHT, "reg": REG, "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                "l2": "inputs (2 CSR orientations 400 MB + factors) exceed the 126 MB L2; no explicit flush",
            },
            "roofline": head["roofline"], "clocks": head["clocks"], "e2e": results.get("e2e"),
            "gpu_launches": launches, "cpu_baseline": cpu,
            "als_fp32": ({"ms_per_epoch": results["fp32"]["ms_per_epoch"], "roofline": results["fp32"]["roofline"]}
                         if "fp32" in results else None),
            "knn": knn,
            "recommend": results.get("recommend"),
        }  # fmt: skip
        emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def bench_recommend(args, tr, dev, peak, peak_src) -> dict:
    """
    SURVEY.md §8f N1/N2: top-100 of all items for a batch of users from the factors just trained —
    one fp32 library GEMM (Q · Xᵀ, item-major) + lk_topn_columns (the reference's heap, one thread
    per user).  CPU figure: the reference's per-query path (f32 matvec + argtopn), one thread.
    """
    import torch

    from lkpy_b200 import engine

    n_top, batch = 100, min(131072, int(tr.d_users.shape[0]))  # one thread per user: large batches fill the GPU
    q, p = tr.d_items, tr.d_users
    rng = np.random.default_rng(11)
    users = torch.from_numpy(np.sort(rng.choice(p.shape[0], batch, replace=False))).to(dev)
    x = p[users].contiguous()
    s = q @ x.T
    engine.topn_columns(s, n_top)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 3
    t_gemm = t_top = 0.0
    for _ in range(reps):
        ev[0].record()
        s = q @ x.T
        ev[1].record()
        idx, val, cnt = engine.topn_columns(s, n_top)
        ev[2].record()
        torch.cuda.synchronize()
        t_gemm += ev[0].elapsed_time(ev[1]) / reps
        t_top += ev[1].elapsed_time(ev[2]) / reps
    n_items = q.shape[0]
    alg = float(n_items) * batch * 4  # the selection reads every score once
    achieved = alg / (t_top * 1e-3) / 1e9
    out = {
        "workload": f"top-{n_top} of {n_items} items for {batch} trained users (ML-25M-shaped ImplicitMF k={K})",
        "users_per_s": batch / ((t_gemm + t_top) * 1e-3), "gemm_ms": t_gemm, "topn_ms": t_top,
        "roofline": {"bound": "hbm", "kernel": "topn_columns_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "algorithmic_bytes": alg,
                     "peak_source": peak_src},
    }  # fmt: skip
    log(f"[bench] recommend: {batch} users x top-{n_top}: GEMM {t_gemm:.1f} ms + top-N {t_top:.1f} ms = "
        f"{out['users_per_s']:.0f} users/s; top-N {achieved:.0f} GB/s = {achieved / peak:.3f} of peak")
    if not args.no_cpu:
        import oracle

        qh, ph = q.cpu().numpy(), x[:256].cpu().numpy()
        t0 = time.perf_counter()
        done = 0
        for u in range(len(ph)):
            sc = qh @ ph[u]
            oracle.argtopn(sc, n_top)
            done += 1
            if time.perf_counter() - t0 > 4.0:
                break
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": done / dt, "unit": "users/s", "cores": 1, "kind": "port",
                               "sample": f"{done} users, f32 matvec + heap argtopn each, one thread"}  # fmt: skip
        # selection parity on the way: the device lists of the first users against the oracle on the device scores
        sh = s[:, :8].cpu().numpy()
        ih = idx[:8].cpu().numpy()
        for c in range(8):
            assert np.array_equal(ih[c], oracle.argtopn(sh[:, c], n_top)), "top-N mismatch vs oracle"
        log(f"[bench] CPU recommend: {out['cpu_baseline']['value']:.0f} users/s (one thread)")
    del s
    return out


def bench_knn(args, inter, dev, peak, peak_src, rank: int = 0, world: int = 1) -> dict | None:
    import torch

    from lkpy_b200 import data, engine

    t0 = time.time()
    kui, kiu, means = data.knn_item_matrices(inter, True)
    log(f"[bench] kNN host prep {time.time() - t0:.1f}s")
    d_ui = engine.DeviceCSR.from_host(kui, dev)
    d_iu = engine.DeviceCSR.from_host(kiu, dev)

    plan = engine.KnnBuildPlan.create(d_ui, d_iu)  # allocates the workspaces once

    if world > 1:
        # item-sharded build: UI replicated, rows dealt by cost, one exchange of the top-K rows
        import torch.distributed as dist

        from lkpy_b200.parallel import sharded_knn_build_topk

        def sbuild():
            plan.prepare()
            return sharded_knn_build_topk(plan, KNN_MIN_SIM, KNN_SAVE)

        sbuild()
        dist.barrier()
        torch.cuda.synchronize()
        reps = max(1, min(args.steps, 3))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            cols, vals, cnt = sbuild()
        s1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([s0.elapsed_time(s1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank != 0:
            return None
        log(f"[bench] kNN build x{world}: {ms:.1f} ms = {inter.n_items / (ms * 1e-3):.0f} items/s")
        return {
            "workload": "ML-25M-shaped synthetic ItemKNNScorer explicit, min_sim=1e-6, save_nbrs=20 (BASELINE configs[2])",
            "build_ms": ms, "build_items_per_s": inter.n_items / (ms * 1e-3),
            "neighbours_kept": int(cnt.sum().item()), "parallelism": f"item rows dealt by cost over {world} GPUs",
        }  # fmt: skip

    def build():
        # the whole build from the resident CSR inputs: tile pointers, row costs, work order,
        # accumulate + top-K, merge, compaction to CSR
        plan.prepare()
        cols, vals, cnt = plan.build_topk(KNN_MIN_SIM, KNN_SAVE)
        return plan, engine.topk_rows_to_csr(cols, vals, cnt)

    plan, csr = build()  # warm-up (also loads the kernels)
    torch.cuda.synchronize()
    reps = max(1, min(args.steps, 3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan, csr = build()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if os.environ.get("LK_BENCH_TRACE"):
        # per-phase wall times with a sync after each phase (diagnostic only)
        def tick(label, fn):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            log(f"[trace] {label}: {(time.perf_counter() - t) * 1e3:.2f} ms")
            return r

        p2 = tick("plan (geometry, tile pointers, cost, argsort)", lambda: engine.KnnBuildPlan.create(d_ui, d_iu))
        r2 = tick("build_topk (accumulate + merge)", lambda: p2.build_topk(KNN_MIN_SIM, KNN_SAVE))
        tick("rows -> CSR", lambda: engine.topk_rows_to_csr(*r2))
    # the accumulate kernel alone (same stream, events around the launch)
    import ctypes as C

    from lkpy_b200 import _lib

    n_items, H = plan.geom.n_items, plan.geom.n_halves
    part_cols = torch.empty(n_items * H * KNN_SAVE, dtype=torch.int32, device=dev)
    part_vals = torch.empty(n_items * H * KNN_SAVE, dtype=torch.float32, device=dev)
    part_cnt = torch.zeros(n_items * H, dtype=torch.int32, device=dev)
    a = plan._args(plan.order, KNN_MIN_SIM, KNN_SAVE)
    a.d_part_cols, a.d_part_vals, a.d_part_cnt = _lib.ptr(part_cols), _lib.ptr(part_vals), _lib.ptr(part_cnt)
    k0.record()
    _lib.check(_lib.lib().lk_knn_build(C.byref(a), _lib.stream_ptr()), "lk_knn_build")
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1)
    cost = plan.cost.cpu().numpy()
    products = int(cost.sum()) - inter.nnz  # sim_row skips the diagonal entry of every (item, user) visit
    nnz_out = int(csr[0][-1].item())
    alg = knn_build_bytes(products, inter.nnz, inter.n_users, inter.n_items, nnz_out)
    achieved = alg / (kms * 1e-3) / 1e9
    out = {
        "workload": "ML-25M-shaped synthetic ItemKNNScorer explicit, min_sim=1e-6, save_nbrs=20 (BASELINE configs[2])",
        "build_ms": ms,
        "build_items_per_s": inter.n_items / (ms * 1e-3),
        "build_kernel_ms": kms,
        "products": products,
        "neighbours_kept": nnz_out,
        "geometry": {"warps": plan.geom.warps, "tile_cols": plan.geom.tile_cols, "halves": plan.geom.n_halves,
                     "ctas_per_sm": plan.geom.ctas_per_sm},
        "roofline": {"bound": "hbm", "kernel": "knn_build_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("knn_build_kernel_bytes", world), "algorithmic_bytes": alg,
                     "peak_source": peak_src},
        "_kui": kui, "_kiu": kiu, "_cost": cost,
    }  # fmt: skip
    log(f"[bench] kNN build: {ms:.1f} ms = {out['build_items_per_s']:.0f} items/s; accumulate kernel {kms:.1f} ms, "
        f"{achieved:.0f} GB/s algorithmic = {achieved / peak:.3f} of peak")

    # scoring: a sample of users against every item
    indptr, c, v = csr
    st = engine.KnnScorerState.create(inter.n_items, indptr, c, v, dev)
    rng = np.random.default_rng(5)
    nq = min(args.score_users, inter.n_users)
    users = np.sort(rng.choice(inter.n_users, nq, replace=False))
    R = inter.coo().tocsr()
    lens = np.diff(R.indptr)[users]
    ref_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.arange(R.indptr[u], R.indptr[u + 1]) for u in users])
    ri = R.indices[idx].astype(np.int32)
    rv = (R.data[idx] - means[ri]).astype(np.float32)
    tgt_ptr = (np.arange(nq + 1, dtype=np.int64) * inter.n_items)
    ti = np.tile(np.arange(inter.n_items, dtype=np.int32), nq)
    d = lambda x: torch.from_numpy(x).to(dev)  # noqa: E731
    d_args = (d(ref_ptr), d(ri), d(rv), d(tgt_ptr), d(ti))
    st.score(*d_args, 20, 1)
    torch.cuda.synchronize()
    e0.record()
    sc, ct = st.score(*d_args, 20, 1)
    e1.record()
    torch.cuda.synchronize()
    sms = e0.elapsed_time(e1)
    out["score"] = {
        "users": nq, "targets_per_user": inter.n_items, "ms": sms, "users_per_s": nq / (sms * 1e-3),
        "scored_fraction": float(torch.isfinite(sc).float().mean().item()),
    }  # fmt: skip
    log(f"[bench] kNN score: {nq} users x all items in {sms:.1f} ms = {out['score']['users_per_s']:.0f} users/s")
    return out


if __name__ == "__main__":
    main()
