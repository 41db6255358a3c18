#!/usr/bin/env python
"""Minimal driver for ncu: two round-trip steps of the bench workload on random data."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from ska_pst_dsp_model_b200 import ops
from ska_pst_dsp_model_b200.filters import load_pst_taps

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
taps = load_pst_taps(normalise=True)
x = torch.view_as_complex(torch.randn((1 << log2n, 2, 2), device="cuda"))
ana = ops.AnalysisPlan(256, "4/3", taps)
inv = ops.InversePlan(256, "4/3", 256, 48, "tukey", deripple_taps=taps)
for _ in range(steps):
    y = inv(ana(x))
torch.cuda.synchronize()
print("ok", tuple(y.shape), ops.launch_count())
