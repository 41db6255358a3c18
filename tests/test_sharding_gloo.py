"""world_size-2 CPU test of the N>1 path: halo exchange over gloo + per-shard compute
(by the oracle, standing in for the kernels) equals the one-shot result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import pfb_oracle as o
    from ska_pst_dsp_model_b200.filters import load_pst_taps
    from ska_pst_dsp_model_b200.sharding import exchange_halo, plan_time_shards
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    taps = load_pst_taps(normalise=True)
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))).astype(np.complex64)   # same on every rank
    sh = plan_time_shards(n, world, 256, "4/3", 3072, 256, 48)[rank]
    local = torch.zeros((sh.own_n + sh.halo_n, 2), dtype=torch.complex64)
    local[: sh.own_n] = torch.from_numpy(x[sh.own0: sh.own0 + sh.own_n])
    exchange_halo(torch.view_as_real(local), sh)
    got = local.numpy()
    ok_halo = np.array_equal(got[: sh.need_n], x[sh.need0: sh.need0 + sh.need_n])
    # per-shard compute; +192 samples because polyphase_analysis.m:62 stops one block short
    seg = got[: min(sh.need_n + 192, got.shape[0])]
    if seg.shape[0] < sh.need_n + 192:
        seg = np.concatenate([seg, np.zeros((sh.need_n + 192 - seg.shape[0], 2), np.complex64)])
    ch = o.polyphase_analysis(np.ascontiguousarray(seg.T[:, None, :]).astype(np.complex128), taps, 256, "4/3")
    y = o.polyphase_synthesis(ch[:, :, : sh.fine_n], True, 256, "4/3", taps, 1, 48, o.window_tukey(256, 48))
    q.put((rank, ok_halo, sh.out0, y[:, 0, :].T.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_halo_exchange_equals_oneshot():
    import torch.multiprocessing as mp
    from oracle import pfb_oracle as o
    from ska_pst_dsp_model_b200.filters import load_pst_taps
    n = 3072 + 192 * (160 * 6 + 96 + 2)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "halo exchange did not reproduce the neighbour's samples"
    taps = load_pst_taps(normalise=True)
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))).astype(np.complex64)
    ch = o.polyphase_analysis(np.ascontiguousarray(x.T[:, None, :]).astype(np.complex128), taps, 256, "4/3")
    ref = o.polyphase_synthesis(ch, True, 256, "4/3", taps, 1, 48, o.window_tukey(256, 48))[:, 0, :].T
    got = np.concatenate([r[3] for r in res], axis=0)
    assert res[1][2] == res[0][3].shape[0]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
