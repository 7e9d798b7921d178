#!/usr/bin/env python
"""Diagnostic: time knn_build_kernel on the ML-25M-shaped matrix for several tilings."""
import ctypes as C
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from lkpy_b200 import _lib, data, engine

inter = data.synth_interactions(**data.ML25M_SHAPE)
kui, kiu, _ = data.knn_item_matrices(inter, True)
dev = _lib.require_device()
d_ui = engine.DeviceCSR.from_host(kui, dev)
d_iu = engine.DeviceCSR.from_host(kiu, dev)
combos = [c.split(":") for c in (sys.argv[1] if len(sys.argv) > 1 else "16:2,8:2,8:3,16:1,32:1,8:4").split(",")]
for warps, ctas in combos:
    _lib.set_option("LK_KNN_WARPS", int(warps))
    _lib.set_option("LK_KNN_CTAS", int(ctas))
    plan = engine.KnnBuildPlan.create(d_ui, d_iu)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cols, vals, cnt = plan.build_topk(1e-6, 20)
        e1.record()
        torch.cuda.synchronize()
    g = plan.geom
    print(f"warps={warps} ctas={ctas} tile_cols={g.tile_cols} halves={g.n_halves} smem={g.smem_bytes}: "
          f"build_topk {e0.elapsed_time(e1):.1f} ms, kept {int(cnt.sum())}", flush=True)
    del plan
    torch.cuda.empty_cache()
