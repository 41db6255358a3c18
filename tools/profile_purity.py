#!/usr/bin/env python
"""Minimal driver for ncu: one chunk of 128 purity vectors of every kind (impulse, tone, comb), no CUDA graph."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from ska_pst_dsp_model_b200.batch import BatchPurity
from ska_pst_dsp_model_b200.filters import load_pst_taps

h = load_pst_taps(normalise=True)
bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=128, use_graph=False)
for _ in range(2):
    rep = bp.sweep(128, 3)
torch.cuda.synchronize()
print("ok", {k: len(v) for k, v in rep.items()})
