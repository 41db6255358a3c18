#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, one rank per GPU): the time-sharded round trip
with NCCL halo exchange must reproduce the single-GPU one-shot result of the whole stream."""
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
rt = ShardedRoundTrip(sh, 256, "4/3", taps, 256, 48, "tukey", True, 2, torch.complex64, dev)
rt.load_owned(x[sh.own0: sh.own0 + sh.own_n])
for _ in range(2):
    y = rt.step()
torch.cuda.synchronize()
ref = ops.InversePlan(256, "4/3", 256, 48, "tukey", deripple_taps=taps, device=dev)(
    ops.AnalysisPlan(256, "4/3", taps, device=dev)(x))
mine = ref[sh.out0: sh.out0 + sh.out_n]
err = float((y - mine).abs().max() / mine.abs().max())
bit = bool(torch.equal(y, mine))
errs = [None] * world
dist.all_gather_object(errs, (rank, err, bit, sh.out_n))
if rank == 0:
    print("sharded-vs-oneshot:", errs, "total out", sum(e[3] for e in errs), "ref", ref.shape[0])
    assert all(e[1] < 1e-6 for e in errs)
    print("OK")
dist.barrier()
dist.destroy_process_group()
