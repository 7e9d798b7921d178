"""
bench_scale.py — the two scale-out configurations of BASELINE.json (configs[3], configs[4];
SURVEY.md §8d "Configs 4 & 5"), run by ``bench.py --workload als100m | knn1b`` under torchrun:

  als100m  100 M-interaction synthetic ImplicitMF, features=128, 1 M users x 200 k items, user/item-sharded
           over the GPUs of one box (rows written into the peer replicas from inside the solve kernel).
           value = epoch time.
  knn1b    1 B-interaction synthetic ItemKNN cosine build (explicit, min_sim=1e-6, save_nbrs=20),
           5 M users x 500 k items (nnz < 2^31: int32 offsets stay valid), item-sharded.
           value = build time -> items/s.

The data are generated, transposed and (kNN) centred / normalised ON THE DEVICE (lkpy_b200.prep: the
same generator construction as the ML-25M-shaped NumPy one, torch's Philox streams) — every rank
generates the same matrix from the same seed, so no interaction ever crosses PCIe or NVLink.  Parity
at scale: a sample of rows is pulled back and compared with the oracle (oracle/parity.py), and the
replicas / results of all ranks are checked for bit equality.
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

from bench import REG, WEIGHT, KNN_MIN_SIM, KNN_SAVE, als_half_bytes, knn_build_bytes, log

SHAPES = {
    "als100m": dict(n_users=1_000_000, n_items=200_000, nnz=100_000_000, k=128),
    "knn1b": dict(n_users=5_000_000, n_items=500_000, nnz=1_000_000_000),
}


def _scaled(shape: dict, s: float) -> dict:
    if s >= 1.0:
        return dict(shape)
    out = dict(shape)
    out["nnz"] = int(shape["nnz"] * s)
    f = s**0.5
    out["n_users"] = max(1000, int(shape["n_users"] * f))
    out["n_items"] = max(500, int(shape["n_items"] * f))
    return out


def _host_rows(m, rows: np.ndarray):
    """Rows ``rows`` of a device CSR as a host InteractionCSR (only those rows travel)."""
    import torch

    from lkpy_b200.data import InteractionCSR

    hp = m.h_indptr.astype(np.int64)
    lens = hp[rows + 1] - hp[rows]
    out_ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = torch.from_numpy(np.concatenate([np.arange(hp[r], hp[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64))
    idx = idx.to(m.indices.device)
    return InteractionCSR(out_ip, m.indices[idx].cpu().numpy(), m.values[idx].cpu().numpy(), (len(rows), m.shape[1]))


def _equal_across_ranks(tensors, dev, world: int) -> bool:
    import torch
    import torch.distributed as dist

    from oracle import parity

    cs = parity.checksum(*[t.cpu().numpy() for t in tensors])
    t = torch.tensor([cs >> 32, cs & 0xFFFFFFFF], dtype=torch.int64, device=dev)
    if world == 1:
        return True
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def _inputs_equal_across_ranks(tensors, dev, world: int) -> bool:
    """Every rank generates the workload itself: check they all generated the same one (integer sums of
    the raw bits, whole and over two strided slices — no large temporaries)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return True
    sig = []
    for t in tensors:
        b = t.view(torch.int32) if t.dtype == torch.float32 else t
        sig += [b.sum(dtype=torch.int64), b[1::2].sum(dtype=torch.int64), b[::3].sum(dtype=torch.int64)]
    sig = torch.stack(sig)
    lo, hi = sig.clone(), sig.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def run(args, rank: int, world: int, dev, peak: float, peak_src: str) -> dict | None:
    return run_als(args, rank, world, dev, peak, peak_src) if args.workload == "als100m" else run_knn(
        args, rank, world, dev, peak, peak_src)  # fmt: skip


# ---------------------------------------------------------------------------
# configs[3]: 100 M interactions, features = 128
# ---------------------------------------------------------------------------


def run_als(args, rank, world, dev, peak, peak_src) -> dict | None:
    import torch

    from lkpy_b200 import engine, prep
    from lkpy_b200.als import ImplicitMFScorer, ImplicitMFTrainer
    from lkpy_b200.components import TrainingOptions

    sh = _scaled(SHAPES["als100m"], args.scale)
    k = sh["k"]

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    u, i, r = prep.synth_interactions_device(sh["n_users"], sh["n_items"], sh["nnz"], device=dev)
    di = prep.DeviceInteractions(u, i, r, sh["n_users"], sh["n_items"])
    if not _inputs_equal_across_ranks([u, i, r], dev, world):
        raise RuntimeError("als100m: the ranks generated different interaction matrices")
    torch.cuda.synchronize()
    log(f"[scale] als100m data on device: {sh['n_users']}x{sh['n_items']}, nnz {di.nnz} ({time.time() - t0:.1f}s)")
    results, parity = {}, {}
    for tag, gdt in (("bf16", "bfloat16"), ("fp32", "float32")):
        if tag not in args.variants.split(","):
            continue
        scorer = ImplicitMFScorer(features=k, epochs=1, regularization=REG, weight=WEIGHT, gather_dtype=gdt)
        if world > 1:
            from lkpy_b200.parallel import ShardedImplicitMFTrainer

            tr = ShardedImplicitMFTrainer(scorer, di, TrainingOptions(rng=42))
        else:
            tr = ImplicitMFTrainer(scorer, di, TrainingOptions(rng=42))
        for _ in range(max(args.warmup, 2)):
            tr.train_epoch_device()
        barrier()
        graphed = bool(world > 1 and tr.enable_graph())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            tr.train_epoch_device()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        tr.kernel_events = []
        for _ in range(min(args.steps, 3)):
            tr.train_epoch_device()
        barrier()
        kern_ms = [a.elapsed_time(b) for a, b in tr.kernel_events]
        n_ev_epochs = min(args.steps, 3)
        tr.kernel_events = None
        tr._raise_on_status()
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        s = 2 if tag == "bf16" else 4
        alg = (als_half_bytes(sh["n_users"], sh["n_items"], di.nnz, k, s, False)
               + als_half_bytes(sh["n_items"], sh["n_users"], di.nnz, k, s, False)) / world  # fmt: skip
        kms = float(np.sum(kern_ms)) / n_ev_epochs
        ach = alg / (kms * 1e-3) / 1e9
        results[tag] = {
            "ms_per_epoch": ms, "solve_kernel_ms_per_epoch": kms, "cuda_graph": graphed,
            "roofline": {"bound": "hbm", "kernel": tr.solve_kernel_name() + " (user + item launch, per GPU)", "achieved": ach,
                         "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "algorithmic_bytes": alg,
                         "peak_source": peak_src},
            "hbm_gb_per_gpu": torch.cuda.max_memory_allocated(dev) / 2**30,
        }  # fmt: skip
        log(f"[scale] als100m {tag} x{world}: {ms:.2f} ms/epoch (solve kernels {kms:.2f} ms), "
            f"{ach:.0f} GB/s algorithmic/GPU = {ach / peak:.3f}; {results[tag]['hbm_gb_per_gpu']:.1f} GB/GPU")
        if not args.no_parity:
            parity["als_" + tag] = _als_parity(tr, di, sh, k, tag, rank, world, dev)
        del tr, scorer
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    head = results.get("bf16") or results["fp32"]
    ok = all(v.get("ok", True) for v in parity.values() if v) if parity else None
    parity["ok"] = ok
    return {
        "metric": "als_implicit_epoch_time_100m_k128", "value": head["ms_per_epoch"], "unit": "ms", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 2), "ms_per_step": head["ms_per_epoch"],
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (bf16-stored gather operand)" if "bf16" in results else "f32", "data": "synthetic (generated on the device)",
        "config": {"workload": "100M-interaction synthetic ImplicitMF features=128, user/item-sharded (BASELINE configs[3])",
                   **{kk: sh[kk] for kk in ("n_users", "n_items")}, "nnz": di.nnz, "features": k, "weight": WEIGHT,
                   "reg": REG, "parallelism": f"row-sharded x{world}, rows peer-written into the replicas by the solve kernel",
                   "scale": args.scale},
        "roofline": head["roofline"], "cuda_graph": head["cuda_graph"], "hbm_gb_per_gpu": head["hbm_gb_per_gpu"],
        "als_fp32": results.get("fp32") if "bf16" in results else None,
        "gpu_launches": 6 * args.steps, "parity": parity if not args.no_parity else None, "e2e": None, "cpu_baseline": None,
    }  # fmt: skip


def _als_parity(tr, di, sh, k, tag, rank, world, dev) -> dict | None:
    """Sampled rows of one user and one item half-step from fixed factors against the f64 oracle; all
    replicas bit-equal across ranks."""
    import torch

    from lkpy_b200 import engine, prep
    from oracle import parity

    bf16 = tag == "bf16"
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    tp0 = torch.randn((sh["n_users"], k), generator=gen, device=dev) * 0.1
    tq0 = torch.randn((sh["n_items"], k), generator=gen, device=dev) * 0.1
    sharded = world > 1

    def half(which):
        tr.d_users.copy_(tp0)
        tr.d_items.copy_(tq0)
        tr.u_plan.status.zero_()
        tr.i_plan.status.zero_()
        if sharded:
            import torch.distributed as dist

            torch.cuda.synchronize()
            dist.barrier()
            tr.half_step(which)
            torch.cuda.synchronize()
            dist.barrier()
        elif which == "user":
            tr._half(tr.u_plan, tr.d_users, tr.d_items, tr.d_items_bf16, tr.config.user_reg)
        else:
            tr._half(tr.i_plan, tr.d_items, tr.d_users, tr.d_users_bf16, tr.config.item_reg)
        torch.cuda.synchronize()
        tr._raise_on_status()
        return (tr.d_users if which == "user" else tr.d_items).clone()

    new_p, new_q = half("user"), half("item")
    eq = _equal_across_ranks([new_p, new_q], dev, world)
    if rank != 0:
        return None
    # the full matrices are not kept by the sharded trainer: rebuild both orientations for the sample
    ui, iu, _ = prep.coo_to_csr_pair(di.users, di.items, tr.prepare_values_device(di), di.n_users, di.n_items)
    out = {"replicas_equal_across_ranks": eq}
    p0, q0 = tp0.cpu().numpy(), tq0.cpu().numpy()
    for which, m, old, other, new in (("user", ui, p0, q0, new_p), ("item", iu, q0, p0, new_q)):
        # rows up to 30 k nonzeros (7 parts): the scalar f64 oracle costs nnz * k^2 per row and all ranks wait for it
        rows = parity.sample_als_rows(m.h_indptr, k, engine.DEFAULT_CHUNK_NNZ, n_random=400, seed=7, max_nnz=30_000,
                                      n_longest=4, n_split=12)
        got = new[torch.from_numpy(rows).to(dev)].cpu().numpy()
        sub = _host_rows(m, rows)
        out[which] = parity.check_als_half("implicit", sub, np.arange(len(rows)), old[rows], other, got, REG, bf16)
    out["ok"] = bool(eq and out["user"]["ok"] and out["item"]["ok"])
    log(f"[scale] parity als100m {tag}: user {out['user']['rel_fro_vs_f64_oracle']:.2e} item "
        f"{out['item']['rel_fro_vs_f64_oracle']:.2e} replicas_equal={eq} ok={out['ok']}")
    del ui, iu
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------
# configs[4]: 1 B interactions, item-kNN build
# ---------------------------------------------------------------------------


def run_knn(args, rank, world, dev, peak, peak_src) -> dict | None:
    import torch

    from lkpy_b200 import engine, prep

    sh = _scaled(SHAPES["knn1b"], args.scale)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    u, i, r = prep.synth_interactions_device(sh["n_users"], sh["n_items"], sh["nnz"], device=dev)
    nnz = int(u.numel())
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    if not _inputs_equal_across_ranks([u, i, r], dev, world):
        raise RuntimeError("knn1b: the ranks generated different interaction matrices")
    t0 = time.time()
    d_ui, d_iu, _means = prep.knn_item_matrices_device(u, i, r, sh["n_users"], sh["n_items"], True)
    del u, i, r
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t_prep = time.time() - t0
    log(f"[scale] knn1b data on device: {sh['n_users']}x{sh['n_items']}, nnz {nnz} (generate {t_gen:.1f}s, "
        f"centre/normalise/transpose on device {t_prep:.1f}s)")
    plan = engine.KnnBuildPlan.create(d_ui, d_iu, world=world)
    products = int(plan.cost.sum().item()) - nnz

    if world > 1:
        from lkpy_b200.parallel import sharded_knn_build_topk

        def build():
            plan.prepare()
            return sharded_knn_build_topk(plan, KNN_MIN_SIM, KNN_SAVE)
    else:
        def build():
            plan.prepare()
            return plan.build_topk(KNN_MIN_SIM, KNN_SAVE)

    cols, vals, cnt = build()  # warm-up
    barrier()
    reps = max(1, min(args.steps, 2))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        cols, vals, cnt = build()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nnz_out = int(cnt.sum().item())
    alg = knn_build_bytes(products, nnz, sh["n_users"], sh["n_items"], nnz_out) / world
    ach = alg / (ms * 1e-3) / 1e9
    mem = torch.cuda.max_memory_allocated(dev) / 2**30
    log(f"[scale] knn1b x{world}: build {ms:.0f} ms = {sh['n_items'] / (ms * 1e-3):.0f} items/s, {products:.3e} products, "
        f"{ach:.0f} GB/s algorithmic/GPU = {ach / peak:.3f}; {mem:.1f} GB/GPU (tile pointers "
        f"{plan.tile_ptr.numel() * 4 / 2**30:.1f} GB)")
    parity = None
    if not args.no_parity:
        parity = _knn_parity(plan, d_ui, d_iu, cols, vals, cnt, rank, world, dev)
    if rank != 0:
        return None
    g = plan.geom
    return {
        "metric": "item_knn_build_items_per_s_1b", "value": sh["n_items"] / (ms * 1e-3), "unit": "items/s", "n_gpus": world,
        "steps": reps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (generated, centred, normalised and transposed on the device)",
        "config": {"workload": "1B-interaction synthetic ItemKNN cosine build, explicit, min_sim=1e-6, save_nbrs=20, "
                               "item-sharded (BASELINE configs[4])",
                   "n_users": sh["n_users"], "n_items": sh["n_items"], "nnz": nnz, "products": products,
                   "parallelism": f"item rows dealt by cost over {world} GPUs, UI replicated, one all-gather of the top-K rows",
                   "geometry": {"warps": g.warps, "tile_cols": g.tile_cols, "halves": g.n_halves, "ctas_per_sm": g.ctas_per_sm},
                   "scale": args.scale},
        "roofline": {"bound": "hbm", "kernel": "knn_build_kernel (whole build, per GPU)", "achieved": ach, "peak": peak,
                     "unit": "GB/s", "frac": ach / peak, "traffic": None, "algorithmic_bytes": alg, "peak_source": peak_src},
        "hbm_gb_per_gpu": mem, "tile_ptr_gb": plan.tile_ptr.numel() * 4 / 2**30, "neighbours_kept": nnz_out,
        "prep_on_device_s": t_prep, "generate_on_device_s": t_gen,
        "gpu_launches": 5 * reps, "parity": parity, "e2e": None, "cpu_baseline": None,
    }  # fmt: skip


def _knn_parity(plan, d_ui, d_iu, cols, vals, cnt, rank, world, dev) -> dict | None:
    """Result bit-equal on all ranks; on rank 0 a sample of item rows (items with at most 4,000 ratings, so
    that the users they touch can be pulled to the host) is compared bit for bit with the oracle's sim_row
    on the sub-matrix of exactly those users."""
    import torch

    from lkpy_b200 import engine
    from lkpy_b200.data import InteractionCSR
    from oracle import parity

    # compare the kept neighbours only: positions past a row's count are uninitialised padding
    mask = torch.arange(cols.shape[1], device=dev)[None, :] < cnt[:, None]
    eq = _equal_across_ranks([torch.where(mask, cols, 0), torch.where(mask, vals, 0.0), cnt], dev, world)
    if rank != 0:
        return None
    n_items = d_iu.shape[0]
    lens = np.diff(d_iu.h_indptr.astype(np.int64))
    rng = np.random.default_rng(3)
    cand = np.flatnonzero((lens > 0) & (lens <= 4000))
    rows = np.sort(rng.choice(cand, min(len(cand), 96), replace=False))
    iu_s = _host_rows(d_iu, rows)
    users = np.unique(iu_s.indices)
    ui_s = _host_rows(d_ui, users)
    # re-index the sampled problem: users -> 0..len(users)-1; columns (items) stay global
    remap = np.full(d_ui.shape[0], -1, dtype=np.int64)
    remap[users] = np.arange(len(users))
    iu_full = InteractionCSR(
        np.zeros(n_items + 1, dtype=np.int64), iu_s.indices, iu_s.values, (n_items, len(users)))  # fmt: skip
    ip = np.zeros(n_items + 1, dtype=np.int64)
    ip[rows + 1] = np.diff(iu_s.indptr)
    iu_full.indptr = np.cumsum(ip)
    iu_full.indices = remap[iu_s.indices].astype(np.int32)
    indptr, c, v = engine.topk_rows_to_csr(cols, vals, cnt)
    res = parity.check_knn_rows(ui_s, iu_full, rows, indptr.cpu().numpy(), c.cpu().numpy(), v.cpu().numpy(),
                                KNN_MIN_SIM, KNN_SAVE)  # fmt: skip
    res["result_equal_across_ranks"] = eq
    res["sample"] = f"{len(rows)} random items with <= 4000 ratings ({len(users)} users touched)"
    res["ok"] = bool(res["ok"] and eq)
    log(f"[scale] parity knn1b: {res['rows']} sampled rows, mismatched {res['n_mismatched']}, "
        f"equal across ranks {eq}, ok={res['ok']}")
    return res
