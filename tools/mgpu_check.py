#!/usr/bin/env python
"""
Multi-GPU check (run under torchrun, one rank per GPU): the row-sharded ALS trainer and the
item-sharded kNN build must reproduce the single-GPU results.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from lkpy_b200 import _lib, data, engine
from lkpy_b200.als import ImplicitMFScorer, ImplicitMFTrainer
from lkpy_b200.components import Dataset, TrainingOptions
from lkpy_b200.parallel import ShardedImplicitMFTrainer, sharded_knn_build_topk

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

inter = data.synth_interactions(20000, 6000, 1_500_000, seed=3)
ds = Dataset(inter)
for gdt in ("float32", "bfloat16"):
    sc = ImplicitMFScorer(features=64, epochs=2, gather_dtype=gdt)
    tr = ShardedImplicitMFTrainer(sc, ds, TrainingOptions(rng=7))
    m = [tr.train_epoch() for _ in range(2)]
    if rank == 0:
        sc1 = ImplicitMFScorer(features=64, epochs=2, gather_dtype=gdt)
        tr1 = ImplicitMFTrainer(sc1, ds, TrainingOptions(rng=7))
        m1 = [tr1.train_epoch() for _ in range(2)]
        eu = np.linalg.norm(sc.user_embeddings - sc1.user_embeddings) / np.linalg.norm(sc1.user_embeddings)
        ei = np.linalg.norm(sc.item_embeddings - sc1.item_embeddings) / np.linalg.norm(sc1.item_embeddings)
        print(f"[{gdt}] sharded x{world} vs single: rel diff users {eu:.2e} items {ei:.2e}; "
              f"deltas {m[-1]} vs {m1[-1]}", flush=True)
        assert eu < 1e-5 and ei < 1e-5
        assert abs(m[-1]["deltaP"] - m1[-1]["deltaP"]) < 1e-3 * abs(m1[-1]["deltaP"]) + 1e-6
    dist.barrier()

kui, kiu, _ = data.knn_item_matrices(inter, True)
plan = engine.KnnBuildPlan.create(
    engine.DeviceCSR.from_host(kui, dev), engine.DeviceCSR.from_host(kiu, dev), world=world
)
cols, vals, cnt = sharded_knn_build_topk(plan, 1e-6, 20)
if rank == 0:
    c1, v1, n1 = plan.build_topk(1e-6, 20)
    torch.cuda.synchronize()
    assert torch.equal(cnt, n1)
    mask = torch.arange(20, device=dev)[None, :] < n1[:, None]
    assert torch.equal(cols[mask], c1[mask]) and torch.equal(vals[mask], v1[mask])
    print(f"kNN sharded x{world}: {int(n1.sum())} neighbours identical to the single-GPU build", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("mgpu check OK", flush=True)
