"""Time-segment sharding of a long stream across the GPUs of one box.

The reference scales stream length only by sequential blocking with tail carry:
``FilterBank.execute`` keeps the unconsumed input tail (FilterBank.m:85-88,119-126)
and ``InverseFilterBank.execute`` the unconsumed fine-channel samples
(InverseFilterBank.m:73-76,104-135).  Those two tails are exactly the halo a
time-sharded run needs where two segments abut:

* overlap-save tail: inversion block ``n`` reads fine samples
  ``[n*keep, n*keep + Nf)``, so the last block of a shard reaches ``2*ov`` fine
  samples into the next shard;
* FIR tail: fine sample ``k`` reads input ``[k*S, k*S + fl)``.

Shards are cut on inversion-block boundaries, rank g owns the input samples from
its first needed sample up to the next rank's first needed sample, and the only
exchange is one neighbour transfer: rank g+1 sends its first
``2*ov*S + fl - S`` samples to rank g (NCCL send/recv over NVLink; gloo on CPU in
the tests).  Every rank then runs the ordinary segment kernels; the concatenated
outputs equal the one-shot result bit for bit (same blocks, same phase index,
same arithmetic).
"""
from __future__ import annotations

from dataclasses import dataclass

from .rational import Rational


@dataclass(frozen=True)
class TimeShard:
    rank: int
    world: int
    inv_block0: int      # first overlap-save block of this rank
    inv_nblocks: int
    fine0: int           # first analysis block (fine-channel sample) needed
    fine_n: int
    need0: int           # first input sample needed
    need_n: int
    own0: int            # input samples this rank holds before the exchange
    own_n: int
    halo_n: int          # samples to fetch from rank+1 (0 for the last rank)
    send_n: int          # samples to send to rank-1 (0 for rank 0)
    out0: int            # first output sample (per pol) of this rank
    out_n: int


def plan_time_shards(n_dat, world, nchan, os_factor, ntaps, input_fft_length, input_overlap, k_total=None):
    """Cut a stream of ``n_dat`` samples/pol into ``world`` time shards for the
    Bunton analysis + inversion round trip."""
    osf = Rational.from_str(os_factor)
    step = nchan * osf.de // osf.nu
    fl = -(-ntaps // nchan) * nchan
    keep = input_fft_length - 2 * input_overlap
    if k_total is None:
        k_total = max((n_dat - fl) // step, 0)      # polyphase_analysis.m:62
    nb_total = max((k_total - 2 * input_overlap) // keep, 0) if k_total >= 2 * input_overlap else 0
    if nb_total < world:
        raise ValueError(f"stream too short: {nb_total} overlap-save blocks for {world} shards")
    okeep = (input_fft_length - 2 * input_overlap) * osf.de // osf.nu * nchan
    base, extra = divmod(nb_total, world)
    starts = [0]
    for g in range(world):
        starts.append(starts[-1] + base + (1 if g < extra else 0))
    shards = []
    for g in range(world):
        n0, nb = starts[g], starts[g + 1] - starts[g]
        fine0 = n0 * keep
        fine_n = (nb - 1) * keep + input_fft_length
        need0 = fine0 * step
        need_n = (fine_n - 1) * step + fl
        if g + 1 < world:
            nxt0 = starts[g + 1] * keep * step
            own_n = nxt0 - need0
            halo_n = need0 + need_n - nxt0
        else:
            own_n = n_dat - need0
            halo_n = 0
        shards.append(dict(rank=g, world=world, inv_block0=n0, inv_nblocks=nb, fine0=fine0, fine_n=fine_n,
                           need0=need0, need_n=need_n, own0=need0, own_n=own_n, halo_n=halo_n,
                           out0=n0 * okeep, out_n=nb * okeep))
    out = []
    for g, s in enumerate(shards):
        send_n = shards[g - 1]["halo_n"] if g > 0 else 0
        if send_n > s["own_n"]:
            raise ValueError("shards are shorter than the halo; use fewer ranks or a longer stream")
        out.append(TimeShard(send_n=send_n, **s))
    return out


def start_halo_exchange(local, shard: TimeShard, group=None):
    """Enqueue the neighbour transfer that fills ``local[own_n : own_n + halo_n]`` with the first
    samples of rank+1; returns the request handles (wait on them before touching the halo).

    ``local``: real-valued view ``[own_n + halo_n, npol, 2]`` (torch.view_as_real
    of the complex shard) on the device the process group communicates on."""
    import torch.distributed as dist
    if shard.world == 1:
        return []
    ops = []
    if shard.send_n > 0:
        ops.append(dist.P2POp(dist.isend, local[: shard.send_n], shard.rank - 1, group))
    if shard.halo_n > 0:
        ops.append(dist.P2POp(dist.irecv, local[shard.own_n: shard.own_n + shard.halo_n], shard.rank + 1, group))
    return dist.batch_isend_irecv(ops)


def exchange_halo(local, shard: TimeShard, group=None):
    for req in start_halo_exchange(local, shard, group):
        req.wait()


def map_peer_buffers(numel_max, dtype, device, group=None):
    """One symmetric buffer per rank (torch.distributed._symmetric_memory: the allocation of every rank is mapped
    into every other rank's address space over NVLink).  Returns (local tensor, [device pointer of every rank's
    buffer]).  Plumbing only - the data path is the analysis kernel reading the neighbour's pointer."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(numel_max, dtype=dtype, device=device)
    hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
    return t, [int(p) for p in hdl.buffer_ptrs], hdl


class ShardedRoundTrip:
    """Analysis + inversion of this rank's time shard (one process per GPU).

    ``halo="peer"`` (default when the process group is NCCL and the symmetric-memory mapping succeeds): the rank's
    input buffer is a symmetric allocation, and the analysis kernel reads the ``halo_n`` samples it needs from rank+1
    straight out of that rank's buffer through a mapped pointer - one analysis launch, no exchange step, no
    collective.  ``halo="exchange"``: the neighbour send/recv of ``start_halo_exchange`` (gloo / CPU tests, and the
    fallback when peer mapping is unavailable)."""

    def __init__(self, shard: TimeShard, nchan, os_factor, taps, input_fft_length, input_overlap,
                 temporal_taper="tukey", deripple=True, npol=2, dtype=None, device=None, halo="auto", group=None,
                 shards=None):
        import torch
        from . import ops
        self.torch = torch
        self.shard = shard
        dtype = dtype or torch.complex64
        self.analysis = ops.AnalysisPlan(nchan, os_factor, taps, "polyphase_analysis", dtype=dtype, device=device)
        self.inverse = ops.InversePlan(nchan, os_factor, input_fft_length, input_overlap, temporal_taper,
                                       deripple_taps=taps if deripple else None, dtype=dtype, device=device)
        dev = self.analysis.device
        self.halo_mode = "exchange"
        self.peer_error = None
        self.next_ptr = 0
        self._symm = None
        if shard.world > 1 and halo in ("auto", "peer"):
            try:
                import torch.distributed as dist
                # identical sizes on every rank: the longest shard (all ranks know every shard's size)
                rows = max((sh.own_n + sh.halo_n) for sh in shards) if shards else None
                if rows is None:
                    t = torch.tensor([shard.own_n + shard.halo_n], dtype=torch.int64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
                    rows = int(t.item())
                buf, ptrs, hdl = map_peer_buffers(rows * npol, dtype, dev, group)
                self._symm = (buf, hdl)
                self.local = buf.view(rows, npol)[: shard.own_n + shard.halo_n]
                self.local.zero_()
                if shard.rank + 1 < shard.world:
                    self.next_ptr = ptrs[shard.rank + 1]
                self.halo_mode = "peer"
            except Exception as e:          # noqa: BLE001  (no P2P / symmetric memory: exchange instead)
                if halo == "peer":
                    raise
                self.peer_error = repr(e)
        if self.halo_mode != "peer":
            self.local = torch.zeros((shard.own_n + shard.halo_n, npol), dtype=dtype, device=dev)
        self.chan = torch.empty((shard.fine_n, nchan, npol), dtype=dtype, device=dev)
        self.out = torch.empty((shard.out_n, npol), dtype=dtype, device=dev)

    def load_owned(self, x_owned, group=None):
        """Copy this rank's samples in; in peer mode every rank must have loaded before any rank runs (one barrier
        here, none per step)."""
        self.local[: self.shard.own_n].copy_(x_owned)
        if self.halo_mode == "peer":
            import torch.distributed as dist
            self.torch.cuda.synchronize()
            dist.barrier(group=group)

    def step(self, group=None):
        self.run_analysis(group)
        self.inverse(self.chan, out=self.out)
        return self.out

    def run_analysis(self, group=None):
        """peer mode: one launch, the halo is read in place from rank+1.  exchange mode: halo exchange overlapped with
        the analysis of the interior blocks (those whose FIR window lies inside the owned samples), then the tail
        blocks."""
        s = self.shard
        if self.halo_mode == "peer":
            self.analysis.segment_peer(self.local[: s.own_n], self.next_ptr, s.halo_n, s.own0, s.fine0, s.fine_n,
                                       out=self.chan)
            return
        reqs = start_halo_exchange(self.torch.view_as_real(self.local), s, group)
        if not reqs:
            self.analysis.segment(self.local, s.own0, s.fine0, s.fine_n, out=self.chan)
        else:
            a = self.analysis
            k_int = max(0, (s.own_n - a.fl) // a.step + 1)
            nu = a.os_factor.nu
            k_int = min(s.fine_n, k_int // nu * nu)        # multiple of nu: the tail launch starts rotation-free
            if k_int > 0:
                a.segment(self.local[: s.own_n], s.own0, s.fine0, k_int, out=self.chan[:k_int])
            for r in reqs:
                r.wait()
            if k_int < s.fine_n:
                a.segment(self.local, s.own0, s.fine0 + k_int, s.fine_n - k_int, out=self.chan[k_int:])
