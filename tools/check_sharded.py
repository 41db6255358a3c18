#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, one rank per GPU): the time-sharded round trip must reproduce the
single-GPU one-shot result of the whole stream bit for bit, in both halo modes: "peer" (the analysis kernel reads the
neighbour's buffer through a mapped pointer over NVLink, no exchange step) and "exchange" (NCCL send/recv)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from ska_pst_dsp_model_b200 import ops
from ska_pst_dsp_model_b200.filters import load_pst_taps
from ska_pst_dsp_model_b200.sharding import ShardedRoundTrip, plan_time_shards

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = world << 21
taps = load_pst_taps(normalise=True)
g = torch.Generator(device="cpu").manual_seed(7)
x = torch.view_as_complex(torch.randn((n, 2, 2), generator=g)).to(dev)          # same stream on every rank
sh = plan_time_shards(n, world, 256, "4/3", 3072, 256, 48)[rank]
shards = plan_time_shards(n, world, 256, "4/3", 3072, 256, 48)
ref = ops.InversePlan(256, "4/3", 256, 48, "tukey", deripple_taps=taps, device=dev)(
    ops.AnalysisPlan(256, "4/3", taps, device=dev)(x))
mine = ref[sh.out0: sh.out0 + sh.out_n]
for mode in os.environ.get("HALO", "peer,exchange").split(","):
    rt = ShardedRoundTrip(sh, 256, "4/3", taps, 256, 48, "tukey", True, 2, torch.complex64, dev, halo=mode, shards=shards)
    rt.load_owned(x[sh.own0: sh.own0 + sh.own_n])
    for _ in range(2):
        y = rt.step()
    torch.cuda.synchronize()
    # timing of the analysis stage alone (device time, max over ranks)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(20):
        rt.run_analysis()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    err = float((y - mine).abs().max() / mine.abs().max())
    bit = bool(torch.equal(y, mine))
    errs = [None] * world
    dist.all_gather_object(errs, (rank, err, bit, sh.out_n, rt.halo_mode))
    if rank == 0:
        print(f"halo={mode}: sharded-vs-oneshot", errs, "total out", sum(e[3] for e in errs), "ref", ref.shape[0],
              f"analysis stage {float(ms):.4f} ms")
        assert all(e[1] < 1e-6 and e[2] for e in errs) and all(e[4] == mode for e in errs)
    del rt
if rank == 0:
    print("OK")
dist.barrier()
dist.destroy_process_group()
