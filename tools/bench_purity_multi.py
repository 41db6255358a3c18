#!/usr/bin/env python
"""BASELINE configs[4]: the batched purity sweep dealt round-robin over the GPUs of one box (replicas, no data-path
collective; the per-vector records are gathered at the end).  Launch with torchrun; rank 0 prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from ska_pst_dsp_model_b200.batch import BatchPurity, gather_reports
from ska_pst_dsp_model_b200.filters import load_pst_taps

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl")
VB = 128
nvec = int(os.environ.get("NVEC", str(VB * 8 * 8)))   # per kind (impulse, tone, comb): whole chunks on 1..8 GPUs
h = load_pst_taps(normalise=True)
bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=VB)
bp.reserve(nvec // world + VB)                      # pinned staging for this rank's share
bp.sweep(VB * world, 3, rank, world)                # warm-up: scratch sized, graphs for the chunk size captured
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
rep = bp.sweep(nvec, 3, rank, world)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.time() - t0
rep = gather_reports(rep, world)
if rank == 0:
    n = sum(len(v) for v in rep.values())
    worst_t = max(r["max_spurious_power"] for r in rep["test_time_domain_impulse"] if 10752 < r["arg"] < 120000)
    worst_f = max(r["max_spurious_power"] for r in rep["test_complex_sinusoid"])
    print(json.dumps({"workload": "batched purity sweep (impulse + tone + comb thirds), 147456-sample vectors, SKA-Low PST round trip",
                      "n_gpus": world, "vectors": n, "seconds": dt, "vectors_per_s": n / dt,
                      "Msamples_s": n * 147456 / dt / 1e6, "worst_max_spurious_dB_time": worst_t,
                      "worst_max_spurious_dB_freq": worst_f}))
if world > 1:
    dist.destroy_process_group()
