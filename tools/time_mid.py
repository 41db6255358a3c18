#!/usr/bin/env python
"""Per-kernel CUDA-event timing of the Mid (cfg3) analysis at 2 pol x 2^24: FIR and row FFT separately (via the C ABI's
own kernels, timed as whole calls with the generic fallbacks switched off)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from ska_pst_dsp_model_b200 import ops
from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter_two_stage

hm = design_pfb_fir_filter_two_stage(4096, 8, 7)
n = 1 << 24
x = torch.view_as_complex(torch.randn((n, 2, 2), device="cuda"))
ana = ops.AnalysisPlan(4096, "8/7", hm, "polyphase_analysis_padded")
inv = ops.InversePlan(4096, "8/7", 512, 128, "tukey", deripple_taps=hm)
ch = ana(x)
y = inv(ch)
torch.cuda.synchronize()


def t(fn, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("analysis_ms %.4f inversion_ms %.4f" % (t(lambda: ana(x, out=ch)), t(lambda: inv(ch, out=y))))
