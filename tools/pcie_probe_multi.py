#!/usr/bin/env python
"""The host-side ceiling of the end-to-end path with ALL GPUs copying at once (torchrun, one rank per GPU): pinned
H2D, D2H and both directions, timed between barriers; rank 0 prints the per-GPU and aggregate rates.  Explains the
e2e scaling of bench.py: the copies of N ranks share the host's memory / root complexes."""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


def timed(fn, reps=6):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    t = torch.tensor([best], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {"n_gpus": world, "bytes_per_copy": n}
for name, fn, mult in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2)):
    ms = timed(fn)
    res[name] = {"ms_slowest_rank": ms, "GBps_per_gpu_per_direction": n / ms / 1e6, "GBps_aggregate": world * mult * n / ms / 1e6}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
