#!/usr/bin/env python
"""Diagnostic: per-phase SM-cycle accounting of the tensor-core ALS kernel on the ML-25M-shaped matrix."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from lkpy_b200 import _lib, data, engine
from lkpy_b200.als import ImplicitMFScorer, ImplicitMFTrainer
from lkpy_b200.components import Dataset, TrainingOptions

inter = data.synth_interactions(**data.ML25M_SHAPE)
sc = ImplicitMFScorer(features=64, epochs=1, gather_dtype="bfloat16")
tr = ImplicitMFTrainer(sc, Dataset(inter), TrainingOptions(rng=42))
for _ in range(3):
    tr.train_epoch_device()
torch.cuda.synchronize()
names = ["fetch", "gather+mma", "wait acc", "readout+fence", "split", "preload (+prefetch)", "end sync", "solve4", "write-back"]
for label, plan, this, other, obf, reg in (
    ("user half", tr.u_plan, tr.d_users, tr.d_items, tr.d_items_bf16, 0.1),
    ("item half", tr.i_plan, tr.d_items, tr.d_users, tr.d_users_bf16, 0.1),
):
    buf = torch.zeros(16, dtype=torch.int64, device=this.device)
    engine.PROF_BUFFER = buf
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    otor = engine.als_otor(other, reg, tr.otor_ws, obf)
    e0.record()
    engine.als_half_epoch(plan, _lib.LK_ALS_IMPLICIT, this, obf, otor=otor)
    e1.record()
    torch.cuda.synchronize()
    engine.PROF_BUFFER = None
    c = buf.cpu().numpy().astype(np.float64)
    tot = c.sum()
    print(f"{label}: {e0.elapsed_time(e1):.3f} ms; chunks {plan.n_chunks}, split rows {plan.n_split_rows}")
    for n, v in zip(names, c[: len(names)]):
        print(f"   {n:14s} {v / tot * 100:5.1f}%   {v / max(plan.n_chunks / 4, 1):10.0f} cycles/group")
