#!/usr/bin/env python
"""Diagnostic: one lk_topn_columns launch (top-100 of 59,047 scores for 32,768 vectors) for an ncu capture."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from lkpy_b200 import _lib, engine

dev = _lib.require_device()
torch.manual_seed(0)
s = torch.randn((59047, 32768), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
engine.topn_columns(s, 100)
torch.cuda.synchronize()
e0.record()
engine.topn_columns(s, 100)
e1.record()
torch.cuda.synchronize()
print(f"topn_columns: {e0.elapsed_time(e1):.2f} ms, {s.numel() * 4 / e0.elapsed_time(e1) / 1e6:.0f} GB/s")
