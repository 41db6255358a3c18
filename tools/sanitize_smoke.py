#!/usr/bin/env python
"""Small shapes through every hand-written producer/consumer pipeline (TMA + mbarrier rings of K1, K2, K3,
pfb_fir_phases) for compute-sanitizer (--tool racecheck / memcheck): few blocks, same code paths as the bench."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from ska_pst_dsp_model_b200 import ops
from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter_two_stage, load_pst_taps

what = sys.argv[1] if len(sys.argv) > 1 else "low"
h = load_pst_taps(normalise=True)
g = torch.Generator(device="cuda").manual_seed(1)
if what == "low":
    n = 3072 + 192 * 16 * 80                       # 1280 blocks: fast K1 (80 CTAs x 16 blocks), K2 TMA ring, K3 ws (8 blocks)
    x = torch.view_as_complex(torch.randn((n, 2, 2), device="cuda", generator=g))
    ana = ops.AnalysisPlan(256, "4/3", h)
    inv = ops.InversePlan(256, "4/3", 256, 48, "tukey", deripple_taps=h)
    ch = ana(x)
    y = inv(ch)
    # two-buffer (peer) form of K1
    cut = n - 21312
    ch2 = ana.segment_peer(x[:cut].clone(), x[cut:].clone().data_ptr(), n - cut, 0, 0, ch.shape[0])
    torch.cuda.synchronize()
    print("low ok", tuple(ch.shape), tuple(y.shape), bool(torch.equal(ch, ch2)), ops.launch_count())
else:
    hm = design_pfb_fir_filter_two_stage(4096, 8, 7)
    n = 1 << 20
    x = torch.view_as_complex(torch.randn((n, 2, 2), device="cuda", generator=g))
    ana = ops.AnalysisPlan(4096, "8/7", hm, "polyphase_analysis_padded")
    ch = ana(x)
    torch.cuda.synchronize()
    print("mid ok", tuple(ch.shape), ops.launch_count())
