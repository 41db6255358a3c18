#!/usr/bin/env python
"""ncu driver: one Mid (cfg3) analysis + inversion on random data."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from ska_pst_dsp_model_b200 import ops
from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter_two_stage

hm = design_pfb_fir_filter_two_stage(4096, 8, 7)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
x = torch.view_as_complex(torch.randn((n, 2, 2), device="cuda"))
ana = ops.AnalysisPlan(4096, "8/7", hm, "polyphase_analysis_padded")
inv = ops.InversePlan(4096, "8/7", 512, 128, "tukey", deripple_taps=hm)
for _ in range(2):
    y = inv(ana(x))
torch.cuda.synchronize()
print("ok", tuple(y.shape))
