#!/usr/bin/env python
"""Where the wall clock of the batched purity sweep goes: host enqueue loop / device wait + read-back / report building."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from ska_pst_dsp_model_b200.batch import BatchPurity
from ska_pst_dsp_model_b200.filters import load_pst_taps

h = load_pst_taps(normalise=True)
bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=128)
bp.sweep(128, 3)
torch.cuda.synchronize()
for rep in range(3):
    bp.timing = {}
    t0 = time.perf_counter()
    r = bp.sweep(1408, 3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nvec = sum(len(v) for v in r.values())
    print(f"{nvec} vectors {dt * 1e3:.2f} ms = {nvec / dt / 1e3:.1f} k vectors/s; per kind [enqueue, wait+readback, report] ms:",
          {k: [round(x * 1e3, 2) for x in v] for k, v in bp.timing.items()})
