#!/usr/bin/env python
"""ncu driver: BASELINE configs[0] (8 channels, OS 8/7, PST taps, 2^18 single-pol samples), analysis + inversion."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from ska_pst_dsp_model_b200 import ops
from ska_pst_dsp_model_b200.filters import load_pst_taps

h = load_pst_taps(normalise=True)
n = 1 << 18
x = torch.view_as_complex(torch.randn((n, 1, 2), device="cuda"))
ana = ops.AnalysisPlan(8, "8/7", h)
inv = ops.InversePlan(8, "8/7", 256, 48, "tukey", deripple_taps=None)
for _ in range(3):
    y = inv(ana(x))
torch.cuda.synchronize()
print("ok", tuple(y.shape))
