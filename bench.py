#!/usr/bin/env python
"""Benchmark of the PFB analysis + inversion hot path (BASELINE.json metric:
input Msamples/s, analysis+inversion, with HBM GB/s vs roofline).

  python bench.py --gpus 1 --steps 10 --warmup 3            # B200 arm
  python bench.py --impl reference --steps 3 --warmup 1     # CPU reference arm
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N ...                  # N > 1 (one rank per GPU)

Workload (configs[1]): SKA-Low PST round trip - 256 channels, OS 4/3,
3072-tap PST filter, polyphase_analysis + polyphase_synthesis inversion (Nf=256,
overlap 48, tukey taper, deripple), dual-pol, 2^24 complex64 samples per pol and
per GPU.  A "step" is one full pass: (halo exchange when N>1) + analysis kernel +
inversion kernel chain over the rank's time shard.  At N>1 the stream is N times
longer and time-sharded (weak scaling); there is no data-path collective: the analysis
kernel reads the neighbour's halo in place through a mapped peer pointer (NCCL send/recv
only as the fallback, ska_pst_dsp_model_b200/sharding.py).

One JSON line is printed by rank 0.  `value` = device-resident throughput over EXACTLY K
steps with nothing but the steps between the two timing events; the per-kernel split
(`kernels`, `roofline`) comes from a second pass of the same K steps with CUDA events between
the kernels (those event records cost ~5 % of a step, `ms_per_step_with_kernel_events`).
`e2e` = the same through the public API with pinned HOST buffers (H2D + D2H inside the timed
region).  `roofline` describes the dominant kernel and carries the SURVEY 8(d) round-trip
figure as `pipeline_frac`.  Side measurements in the same run: `sustained` (>= 1 s of steps,
clocks and power), `fused_round_trip` (2-kernel in-process form, own key), `e2e_int16`,
`other_configs` (BASELINE configs[0], [2], [4] and the fp64 mode), `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NCHAN, OS, NFFT, OVERLAP, NPOL = 256, "4/3", 256, 48, 2
METRIC = "input Msamples/s (PFB analysis + inversion round trip)"
UNIT = "Msamples/s"


def workload_config(log2n, world, halo=None):
    return {
        "workload": f"SKA-Low PST round trip: {NCHAN} ch, OS {OS}, 3072-tap PST filter, Nf={NFFT}, overlap={OVERLAP}, "
                    f"tukey+deripple, {NPOL} pol x 2^{log2n} complex64 samples per GPU",
        "channels": NCHAN, "os_factor": OS, "fir_taps": 3072, "input_fft_length": NFFT, "input_overlap": OVERLAP,
        "n_pol": NPOL, "samples_per_pol_per_gpu": 1 << log2n, "global_samples_per_pol": world << log2n,
        "parallelism": f"time-shard x{world}" if world > 1 else "single GPU",
        "halo": halo,
        "l2": "inputs (256 MiB/GPU) and outputs larger than the 126 MB L2; no flush needed",
    }


# --------------------------------------------------------------------------
# synthetic stream: 32-tone comb (sgcht.m:394-432) + an impulse every output_keep
# samples, a pure function of the absolute sample index (any shard can be
# regenerated; no RNG)
# --------------------------------------------------------------------------
def comb_params():
    import numpy as np
    amps = np.linspace(1.0, 2.0 ** 0.5, 32)
    fmin = -0.5 + 1.0 / 128 + 1.0 / (NCHAN * 4)
    freqs = np.linspace(fmin, fmin + 31.0 / 32, 32)
    return amps, freqs


def comb_period_table():
    """The comb's 32 tones sit 1/32 cycles/sample apart, so x[t] = exp(2 pi i f0 t) * P[t mod 32] / 32 with
    P[m] = sum_k a_k exp(2 pi i k m / 32): one complex exponential per sample instead of 32."""
    import numpy as np
    amps, freqs = comb_params()
    m = np.arange(32)
    P = (amps[None, :] * np.exp(2j * np.pi * ((np.arange(32)[None, :] * m[:, None]) % 32) / 32.0)).sum(axis=1)
    return float(freqs[0]), P / 32.0


def synth_numpy(t0, n, npol):
    import numpy as np
    f0, P = comb_period_table()
    out = np.empty((n, npol), dtype=np.complex64)
    chunk = 1 << 21
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        ti = np.arange(t0 + c0, t0 + c0 + m, dtype=np.int64)
        x = np.exp(2j * np.pi * np.mod(f0 * ti.astype(np.float64), 1.0)) * P[ti % 32]
        x[(ti % 30720) == 0] += 1.0
        out[c0:c0 + m] = x.astype(np.complex64)[:, None]
    return out


def synth_torch(t0, n, npol, device):
    import torch
    f0, P = comb_period_table()
    Pd = torch.from_numpy(P).to(device)
    out = torch.empty((n, npol), dtype=torch.complex64, device=device)
    chunk = 1 << 22
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        ti = torch.arange(t0 + c0, t0 + c0 + m, dtype=torch.int64, device=device)
        ph = torch.remainder(ti.to(torch.float64) * f0, 1.0) * (2 * torch.pi)
        acc = torch.polar(torch.ones_like(ph), ph) * Pd[ti % 32]
        acc[(ti % 30720) == 0] += 1.0
        out[c0:c0 + m] = acc.to(torch.complex64)[:, None]
    return out


# --------------------------------------------------------------------------
# CPU reference arm: the numpy oracle (port of the Matlab model) on all host cores
# --------------------------------------------------------------------------
_CPU = {}
_THREAD_VARS = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "VECLIB_MAXIMUM_THREADS")


def host_cores():
    """Cores this process may run on (the cgroup / affinity mask, not the machine's count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def _pin_blas_threads():
    """One BLAS/OpenMP thread per worker process: the arm parallelises over processes, and unpinned BLAS pools
    oversubscribe the cores (round 1: 3.3 vs 15.6 Msamples/s on the same 8 cores)."""
    for k in _THREAD_VARS:
        os.environ[k] = "1"


def _cpu_init():
    _pin_blas_threads()
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limits"] = threadpool_limits(limits=1)
    except Exception:           # noqa: BLE001
        pass


def _cpu_worker(job):
    import numpy as np
    from oracle import pfb_oracle as orc
    seg_in0, seg_n, out0, out_n = job
    x = _CPU["x"][seg_in0: seg_in0 + seg_n]                      # [t][pol]
    xo = np.ascontiguousarray(x.T[:, None, :])
    ch = orc.polyphase_analysis(xo, _CPU["taps"], NCHAN, OS)
    y = orc.polyphase_synthesis(ch, True, NFFT, OS, _CPU["taps"], 1, OVERLAP, _CPU["taper"])
    out = np.frombuffer(_CPU["out"], dtype=np.complex64).reshape(-1, NPOL)
    out[out0: out0 + out_n] = y[:, 0, :out_n].T
    return out_n


def cpu_round_trip_setup(log2n, workers):
    import multiprocessing as mp
    import numpy as np
    from oracle import pfb_oracle as orc
    from ska_pst_dsp_model_b200.filters import load_pst_taps
    from ska_pst_dsp_model_b200.sharding import plan_time_shards
    n = 1 << log2n
    nseg = workers * 4
    while True:                      # short samples: fewer segments than 4 per worker
        try:
            shards = plan_time_shards(n, nseg, NCHAN, OS, 3072, NFFT, OVERLAP)
            break
        except ValueError:
            if nseg == 1:
                raise
            nseg //= 2
    _CPU["x"] = synth_numpy(0, n, NPOL)
    _CPU["taps"] = load_pst_taps(normalise=True)
    _CPU["taper"] = orc.window_tukey(NFFT, OVERLAP)
    total_out = shards[-1].out0 + shards[-1].out_n
    _CPU["out"] = mp.RawArray("f", total_out * NPOL * 2)
    # the reference's nblocks = floor((n_dat - fl)/step) (polyphase_analysis.m:62) is one
    # short of what fits, so hand each segment one extra step of input
    jobs = [(s.need0, s.need_n + 192, s.out0, s.out_n) for s in shards]
    _pin_blas_threads()
    pool = mp.get_context("fork").Pool(workers, initializer=_cpu_init)
    return pool, jobs, n * NPOL


def run_reference(args):
    """`--impl reference`: the oracle port on the host cores.  One step = the whole per-GPU
    configuration (2^24 samples/pol x 2 pol) unless --cpu-log2-samples bounds it; at N > 1 the stream of the
    B200 arm is N times longer, the CPU step stays one GPU's share (a bounded sample: the metric is a rate)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    _pin_blas_threads()
    workers = min(host_cores(), 64)
    log2n = args.cpu_log2_samples if args.cpu_log2_samples else args.log2_samples
    pool, jobs, samples = cpu_round_trip_setup(log2n, workers)
    for _ in range(args.warmup):
        pool.map(_cpu_worker, jobs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.map(_cpu_worker, jobs)
    dt = (time.perf_counter() - t0) / args.steps
    pool.close()
    value = samples / dt / 1e6
    whole = log2n == args.log2_samples
    sample = (f"{'the whole' if whole else 'first'} 2^{log2n} samples/pol x {NPOL} pol of the same synthetic stream per step "
              f"({'= one GPU share of the config' if whole else 'bounded prefix'}), {len(jobs)} time segments over "
              f"{workers} processes, 1 BLAS/OpenMP thread each")
    cfg = workload_config(args.log2_samples, args.gpus)
    cfg["cpu_step"] = sample
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "io_dtype": "complex64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "threads_per_process": 1, "host_cpu_count": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "numpy port of the reference's Matlab model (oracle/pfb_oracle.py), fp64 arithmetic as the Matlab scratch "
                "arrays (polyphase_synthesis.m:152-154); the reference itself cannot run here (Matlab / un-vendored pfb package)",
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled through NVML from a thread every ~1 ms
    (nvidia-smi -lms cannot resolve a timed region of a few milliseconds)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.h = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception as e:          # noqa: BLE001
            self.h = None
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), clk, rs, pw))
            except Exception:           # noqa: BLE001
                pass
            time.sleep(0.001)

    def stop(self, t_begin, t_end):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        inside = [s for s in self.samples if t_begin <= s[0] <= t_end] or self.samples[-3:]
        reasons = set()
        for _, _, rs, _ in inside:
            for nm, bit in names.items():
                if rs & bit:
                    reasons.add(nm)
        clks = [s[1] for s in inside]
        return {"sm_mhz": statistics.median(clks) if clks else None, "sm_max_mhz": self.smax,
                "reasons": sorted(reasons), "samples": len(inside),
                "power_w_max": max(s[3] for s in inside) if inside else None}


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f)
        except Exception:
            pass
    return {}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from ska_pst_dsp_model_b200 import ops
    from ska_pst_dsp_model_b200.filters import load_pst_taps
    from ska_pst_dsp_model_b200.sharding import ShardedRoundTrip, plan_time_shards

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    log2n = args.log2_samples
    n_global = world << log2n
    taps = load_pst_taps(normalise=True)
    shards = plan_time_shards(n_global, world, NCHAN, OS, 3072, NFFT, OVERLAP)
    shard = shards[rank]
    rt = ShardedRoundTrip(shard, NCHAN, OS, taps, NFFT, OVERLAP, "tukey", True, NPOL, torch.complex64, dev,
                          halo=args.halo, shards=shards)
    if args.scratch_mib:
        rt.inverse.set_scratch_bytes(args.scratch_mib << 20)
    rt.load_owned(synth_torch(shard.own0, shard.own_n, NPOL, dev))
    samples_rank = shard.own_n * NPOL
    osf = 4.0 / 3.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        rt.step()
    barrier()

    # ---- timed region: EXACTLY K steps, device-resident input, nothing but the steps between the two events -------
    K = args.steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    launches0 = ops.launch_count()
    barrier()
    t_begin = time.perf_counter()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(K):
        rt.run_analysis()
        rt.inverse(rt.chan, out=rt.out)
    end.record()
    barrier()
    t_end = time.perf_counter()
    launches = ops.launch_count() - launches0
    ms_total = start.elapsed_time(end)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms_step = ms_total / K
    t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(samples_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step = float(t.item())
    total_samples = float(tot.item())
    value = total_samples / (ms_step * 1e-3) / 1e6

    # ---- per-kernel split: the same K steps once more with CUDA events between the kernels (the event records
    #      themselves cost ~5 % of a step - measured: 0.468 vs 0.444 ms - so they stay out of the region `value` is
    #      computed from; `ms_per_step_with_kernel_events` is this pass)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    rt.inverse.set_profiling(True)
    barrier()
    s0e, e0e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0e.record()
    for i in range(K):
        ev[i][0].record()
        rt.run_analysis()
        ev[i][1].record()
        rt.inverse(rt.chan, out=rt.out)
        ev[i][2].record()
    e0e.record()
    barrier()
    ms_with_events = s0e.elapsed_time(e0e) / K
    prof = rt.inverse.get_profile()
    rt.inverse.set_profiling(False)

    ms_analysis = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    ms_inverse = sum(e[1].elapsed_time(e[2]) for e in ev) / K

    # ---- sustained: the same step back to back for >= --min-seconds (clocks / power under a long run) ----
    sustained = None
    if args.min_seconds > 0:
        ks = max(K, int(args.min_seconds * 1e3 / ms_step) + 1)
        samp2 = ClockSampler(local_rank)
        if rank == 0:
            samp2.start()
        barrier()
        tb = time.perf_counter()
        s1, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s1.record()
        for _ in range(ks):
            rt.run_analysis()
            rt.inverse(rt.chan, out=rt.out)
        e1.record()
        barrier()
        te = time.perf_counter()
        ms_s = s1.elapsed_time(e1) / ks
        ts = torch.tensor([ms_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_s = float(ts.item())
        launches += 0   # (not part of the K-step claim)
        sustained = {"steps": ks, "seconds": ms_s * ks * 1e-3, "ms_per_step": ms_s,
                     "value": total_samples / (ms_s * 1e-3) / 1e6, "unit": UNIT,
                     "vs_burst": ms_step / ms_s, "clocks": samp2.stop(tb, te) if rank == 0 else None}

    # ---- per-kernel roofline (algorithmic bytes = the kernel's compulsory I/O) -
    cb = 8  # complex64
    bytes_analysis = (shard.need_n * NPOL + shard.fine_n * NCHAN * NPOL) * cb
    bytes_chan = 2 * shard.fine_n * NCHAN * NPOL * cb
    bytes_resample = (shard.fine_n * NCHAN * NPOL + shard.out_n * NPOL) * cb
    n0, n1 = max(prof["chan_idft"][1], 1), max(prof["resample"][1], 1)
    kernels = {
        "pfb_analysis": {"ms_per_step": ms_analysis, "launches_per_step": 1, "bytes_per_step": bytes_analysis},
        "pfb_chan_idft": {"ms_per_step": prof["chan_idft"][0] / K, "launches_per_step": n0 / K, "bytes_per_step": bytes_chan},
        "pfb_resample": {"ms_per_step": prof["resample"][0] / K, "launches_per_step": n1 / K, "bytes_per_step": bytes_resample},
    }
    peak, peak_src = load_peak()
    for kk in kernels.values():
        kk["gbs"] = kk["bytes_per_step"] / (kk["ms_per_step"] * 1e-3) / 1e9 if kk["ms_per_step"] > 0 else None
        kk["frac"] = kk["gbs"] / peak if kk["gbs"] else None
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    dk = kernels[dom]
    traffic = load_traffic().get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dk["gbs"], "peak": peak, "unit": "GB/s",
                "frac": dk["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dk["bytes_per_step"] / dk["launches_per_step"],
                "avg_launch_ms": dk["ms_per_step"] / dk["launches_per_step"]}
    # the round trip with the channelized data materialised moves 16(1+OS) B per input sample per pol
    # (actual elements counted: the analysis reads need_n, the inversion writes out_n)
    rt_bytes = bytes_analysis + (shard.fine_n * NCHAN * NPOL + shard.out_n * NPOL) * cb
    pipeline = {"algorithmic_bytes_per_step": rt_bytes, "gbs": rt_bytes / (ms_step * 1e-3) / 1e9,
                "frac": rt_bytes / (ms_step * 1e-3) / 1e9 / peak, "ms_analysis": ms_analysis, "ms_inversion": ms_inverse,
                "frac_analysis": bytes_analysis / (ms_analysis * 1e-3) / 1e9 / peak,
                "frac_inversion": (rt_bytes - bytes_analysis) / (ms_inverse * 1e-3) / 1e9 / peak,
                "nominal_16_1plusOS_bytes": 16 * (1 + osf) * samples_rank}
    roofline.update({"pipeline_frac": pipeline["frac"], "pipeline_achieved": pipeline["gbs"],
                     "pipeline_algorithmic_bytes_per_step": rt_bytes,
                     "peak_nominal": 8000.0, "frac_of_nominal": dk["gbs"] / 8000.0,
                     "pipeline_frac_of_nominal": pipeline["gbs"] / 8000.0,
                     "note": "per-kernel times come from a second pass of the same K steps with CUDA events between the kernels; "
                             "frac = dominant kernel's own compulsory bytes / its CUDA-event time / measured copy peak; "
                             "pipeline_frac = SURVEY 8(d) round-trip bytes (16(1+OS) B/sample) / step time / the same peak"})

    # ---- fused analysis -> inversion (own key, never `value`): for in-process pipelines that do not need the channelized
    #      data as a FILE, the analysis kernel writes the inversion's intermediate itself and the channel-IDFT kernel is
    #      not run (ops.round_trip).  The drop-in synthesize() reads channelized files and cannot use this.
    fused = None
    if world == 1 and not args.no_fused:
        try:
            fused = {}
            x_all = rt.local[: shard.own_n]                      # the whole 2^24-sample stream of this GPU
            nb_all = rt.analysis.nblocks(x_all.shape[0])
            assert rt.inverse.out_len(nb_all) == rt.out.shape[0]
            ch_all = torch.empty((nb_all, NCHAN, NPOL), dtype=torch.complex64, device=dev)
            for nm, want in (("with_channelized_output", True), ("inversion_only_output", False)):
                ch_buf = ch_all if want else None
                ops.round_trip(rt.analysis, rt.inverse, x_all, chan=ch_buf, out=rt.out)
                torch.cuda.synchronize()
                s4, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s4.record()
                for _ in range(K):
                    ops.round_trip(rt.analysis, rt.inverse, x_all, chan=ch_buf, out=rt.out)
                e4.record()
                torch.cuda.synchronize()
                ms_f = s4.elapsed_time(e4) / K
                fused[nm] = {"ms_per_step": ms_f, "value": samples_rank / (ms_f * 1e-3) / 1e6, "unit": UNIT,
                             "vs_three_kernel_step": ms_step / ms_f}
            fused["note"] = ("analysis writes N x its rotated arm sums = the inversion's channel-IDFT intermediate; 2 kernels "
                             "instead of 3; in-process pipelines only (purity sweep, host round trip), not the file drop-in")
            del ch_all
            # leave rt.out / rt.chan as the three-kernel step produced them for the checks below
            rt.run_analysis()
            rt.inverse(rt.chan, out=rt.out)
            torch.cuda.synchronize()
        except Exception as e:          # noqa: BLE001
            fused = {"error": repr(e)}

    # ---- e2e: pinned host buffers through the public streaming operator ------
    # (H2D of the rank's samples incl. halo, kernels, D2H of the inverted series, overlapped
    #  over time segments on three streams; the host buffer already holds every halo)
    from ska_pst_dsp_model_b200.streaming import HostRoundTrip
    Ke = max(1, min(K, args.e2e_steps))
    x_host = torch.empty((shard.need_n, NPOL), dtype=torch.complex64).pin_memory()
    if world > 1:           # (setup, untimed) the host buffer holds the halo too: fetch it once, whatever the device halo mode
        from ska_pst_dsp_model_b200.sharding import exchange_halo
        exchange_halo(torch.view_as_real(rt.local), shard)
        torch.cuda.synchronize()
    x_host.copy_(rt.local[: shard.need_n])
    del rt.chan
    hrt = HostRoundTrip(shard.need_n, NCHAN, OS, taps, NFFT, OVERLAP, "tukey", True, NPOL, nseg=args.e2e_segments,
                        device=dev, k_total=shard.fine_n)
    assert hrt.n_out == shard.out_n
    y_host = torch.empty((hrt.n_out, NPOL), dtype=torch.complex64).pin_memory()
    hrt.run(x_host, y_host)
    barrier()
    e2e_match = bool(torch.equal(y_host[:65536], rt.out[:65536].cpu()))
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(Ke):
        hrt.run(x_host, y_host)
    e2.record()
    barrier()
    ms_e2e = s2.elapsed_time(e2) / Ke
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    h2d = sum(sg.need_n for sg in hrt.segs) * NPOL * cb
    e2e = {"value": total_samples / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": hrt.n_out * NPOL * cb, "segments": len(hrt.segs),
           "matches_device_resident_result": e2e_match,
           "api": "streaming.HostRoundTrip.run(x_pinned, y_pinned): ops.AnalysisPlan.segment + ops.InversePlan over "
                  "the C ABI, H2D/kernels/D2H overlapped per time segment"}

    # ---- the same with integer host buffers (NBIT 16 payloads in and out, DADARead.m:41-53): a separate key, the
    #      headline e2e above stays the complex64 configuration BASELINE.json names
    e2e_int16 = None
    if not args.no_int16:
        del hrt
        hi = HostRoundTrip(shard.need_n, NCHAN, OS, taps, NFFT, OVERLAP, "tukey", True, NPOL, nseg=args.e2e_segments,
                           device=dev, k_total=shard.fine_n, in_nbit=16, out_nbit=16, out_scale=1.0)
        xi_host = torch.empty((shard.need_n, NPOL, 2), dtype=torch.int16).pin_memory()
        xi_host.copy_(torch.view_as_real(x_host).mul(8192.0).round_().to(torch.int16))
        yi_host = torch.empty((hi.n_out, NPOL, 2), dtype=torch.int16).pin_memory()
        hi.run(xi_host, yi_host)
        barrier()
        s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s3.record()
        for _ in range(Ke):
            hi.run(xi_host, yi_host)
        e3.record()
        barrier()
        ms_i = s3.elapsed_time(e3) / Ke
        t = torch.tensor([ms_i], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_i = float(t.item())
        # samples scaled by 8192 and rounded to int16; unity-gain round trip, so the int16 output / 8192 agrees with the
        # float path to the two quantisation steps (~1e-4)
        q = (torch.view_as_real(y_host[:65536]) - yi_host[:65536].to(torch.float32) / 8192.0).abs().max()
        e2e_int16 = {"value": total_samples / (ms_i * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_i,
                     "h2d_bytes_per_step": sum(sg.need_n for sg in hi.segs) * NPOL * 4,
                     "d2h_bytes_per_step": hi.n_out * NPOL * 4,
                     "max_abs_diff_vs_float_path": float(q),
                     "note": "int16 (re, im) host buffers both ways (NBIT 16 DADA payloads), cast on the device; "
                             "half the PCIe bytes of the complex64 e2e"}
        hrt = hi
    # quick in-bench sanity check of the result against the input (not the oracle):
    # the round trip reproduces the delayed input (unity gain, deripple on)
    y = rt.out[:4096, 0]
    delay = 3072 // 2 + 36 * 256            # chop.m:36-37 with (dir,kludge) = (-1, 1)
    xin = rt.local[delay: delay + 4096, 0] if shard.rank == 0 else None
    rt_err = float((y - xin).abs().max() / xin.abs().max()) if xin is not None else None

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = min(host_cores(), 64)
        clog2 = args.cpu_log2_samples or min(22, log2n)
        pool, jobs, samples = cpu_round_trip_setup(clog2, workers)
        pool.map(_cpu_worker, jobs[: workers])               # warm the pool
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or time.perf_counter() - t0 < 5.0:
            pool.map(_cpu_worker, jobs)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        pool.close()
        # parity spot check of the GPU result against the oracle on the common prefix
        ref = np.frombuffer(_CPU["out"], dtype=np.complex64).reshape(-1, NPOL)
        m = min(ref.shape[0], 65536)
        got = rt.out[:m].cpu().numpy()
        perr = float(np.abs(got - ref[:m]).max() / np.abs(ref[:m]).max())
        cpu_baseline = {"value": samples / dt / 1e6, "unit": UNIT, "cores": workers, "kind": "port",
                        "sample": f"first 2^{clog2} samples/pol x {NPOL} pol of the same stream, "
                                  f"{reps} repetitions, {len(jobs)} segments over {workers} processes, 1 BLAS thread each",
                        "threads_per_process": 1, "host_cpu_count": os.cpu_count(),
                        "parity_vs_gpu_prefix": perr}

    other = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        del hrt, x_host, y_host
        rt.out = rt.local = None
        torch.cuda.empty_cache()
        try:
            other = measure_other_configs(dev, peak, cpu=not args.no_cpu_baseline)
        except Exception as e:          # noqa: BLE001  (side measurements must not lose the bench line)
            other = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(log2n, world, rt.halo_mode if world > 1 else None), "clocks": clocks, "e2e": e2e,
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels,
        "pipeline": pipeline, "ms_per_step_with_kernel_events": ms_with_events, "sustained": sustained, "fused_round_trip": fused, "e2e_int16": e2e_int16, "round_trip_max_err": rt_err,
        "other_configs": other,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------
# the other BASELINE configs, measured in the same run (parity-test cases, not the bench line)
# --------------------------------------------------------------------------
def _time_cuda(fn, reps):
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _time_cuda_graph(fn, reps):
    """Device time per call with the host out of the loop: `reps` calls captured into one CUDA graph (a 20 us kernel
    cannot be timed through a Python launch loop that costs as much per iteration)."""
    import torch
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
        for _ in range(reps):
            fn()
    torch.cuda.current_stream().wait_stream(cap)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _cpu_cfg_worker(job):
    """One independent stream per worker process through the oracle (analysis + inversion)."""
    import numpy as np
    from oracle import pfb_oracle as orc
    kind, seed = job
    c = _CPU["cfg"][kind]
    rng = np.random.default_rng(seed)
    if kind == "cfg1":
        x = np.zeros((1, 1, c["n"]), dtype=np.complex128)
        x[0, 0, int(0.11 * c["n"])] = 1.0
        ch = orc.polyphase_analysis(x, c["taps"], 8, "8/7")
        y = orc.polyphase_synthesis(ch, True, 256, "8/7", c["taps"], 1, 48, c["taper"])
    else:
        x = rng.standard_normal((1, 1, c["n"])) + 1j * rng.standard_normal((1, 1, c["n"]))
        ch = orc.polyphase_analysis_padded(x, c["taps"], 4096, "8/7")
        y = orc.polyphase_synthesis(ch, True, 512, "8/7", c["taps"], 1, 128, c["taper"])
    return int(y.shape[2])


def _cpu_cfg_rate(kind, cfg, workers):
    """All-core CPU rate of the oracle on `kind`: every worker runs one independent stream of cfg['n'] samples."""
    import multiprocessing as mp
    _CPU.setdefault("cfg", {})[kind] = cfg
    _pin_blas_threads()
    pool = mp.get_context("fork").Pool(workers, initializer=_cpu_init)
    t0 = time.perf_counter()
    pool.map(_cpu_cfg_worker, [(kind, i) for i in range(workers)])
    dt = time.perf_counter() - t0
    pool.close()
    return {"value": workers * cfg["n"] / dt / 1e6, "unit": UNIT, "cores": workers, "kind": "port", "threads_per_process": 1,
            "sample": f"{workers} independent streams of {cfg['n']} samples x 1 pol, one per process, {dt:.1f} s"}


def _cpu_purity_worker(job):
    """One purity test vector through the oracle pipeline + the report's reductions (purity.py:165-263)."""
    import numpy as np
    from oracle import pfb_oracle as orc
    kind, arg = job
    c = _CPU["purity"]
    n = c["n"]
    if kind == 0:
        x = np.zeros(n, dtype=np.complex64)
        x[int(arg)] = 1.0
    else:
        x = np.exp(1j * (2 * np.pi * np.mod(arg * np.arange(n) / n, 1.0) + np.pi / 4)).astype(np.complex64)
    ch = orc.polyphase_analysis(x[None, None, :], c["taps"], NCHAN, OS)
    y = orc.polyphase_synthesis(ch, True, NFFT, OS, c["taps"], 1, OVERLAP, c["taper"])[0, 0]
    a = x[c["shift"]: c["shift"] + y.size]
    d = np.abs(y[: a.size] - a).sum()
    p = np.abs(np.fft.fft(y[: c["fft"]]) / c["fft"]) ** 2
    p[np.argmax(p)] = 0
    return float(d + p.max())


def _cpu_purity_rate(taps, workers):
    import multiprocessing as mp
    from oracle import pfb_oracle as orc
    n = 192 * NCHAN * 3
    _CPU["purity"] = {"n": n, "taps": taps, "taper": orc.window_tukey(NFFT, OVERLAP), "shift": 36 * NCHAN + 1535, "fft": 2 * 192 * NCHAN}
    _pin_blas_threads()
    pool = mp.get_context("fork").Pool(workers, initializer=_cpu_init)
    jobs = [(i & 1, 1000 + 997 * i) for i in range(workers * 6)]
    t0 = time.perf_counter()
    pool.map(_cpu_purity_worker, jobs)
    dt = time.perf_counter() - t0
    pool.close()
    return {"value": len(jobs) * n / dt / 1e6, "unit": UNIT, "vectors_per_s": len(jobs) / dt, "cores": workers, "kind": "port",
            "threads_per_process": 1, "sample": f"{len(jobs)} vectors (impulse / tone) of {n} samples through the oracle "
                                                f"pipeline + report reductions, {dt:.1f} s"}


def measure_other_configs(dev, peak, cpu=True):
    """BASELINE configs[0] (cfg1), configs[2] (cfg3, SKA-Mid at 2^24 x 2 pol) and configs[4] (cfg5, purity sweep)
    on this GPU, each with its own roofline figure and an all-core oracle rate on a bounded sample."""
    import numpy as np
    import torch
    from oracle import pfb_oracle as orc
    from ska_pst_dsp_model_b200 import ops
    from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter_two_stage, load_pst_taps
    workers = min(host_cores(), 64)
    res = {}
    h = load_pst_taps(normalise=True)
    cb = 8
    # ---- cfg1: 8 channels, OS 8/7, 3072 taps (384 per arm: FMA-bound), 2^18 single-pol impulse -------------
    n = 1 << 18
    x = torch.zeros((n, 1), dtype=torch.complex64, device=dev)
    x[int(0.11 * n)] = 1
    ana = ops.AnalysisPlan(8, "8/7", h, device=dev)
    inv = ops.InversePlan(8, "8/7", 256, 48, "tukey", deripple_taps=h, device=dev)
    ch = ana(x)
    y = inv(ch)
    ms_a = _time_cuda_graph(lambda: ana(x, out=ch), 20)
    ms_i = _time_cuda_graph(lambda: inv(ch, out=y), 20)
    flop_a = (4.0 * 3072 / 7 + 5.0 * 8 * 3 / 7 + 6.0 * 8 / 7) * n            # 1759 flop/sample
    by = (n + 2 * ch.numel() + y.numel()) * cb
    res["cfg1"] = {"workload": "8 ch, OS 8/7, PST taps (384/arm), 2^18 x 1 pol impulse; inversion Nf=256 ov=48",
                   "ms_analysis": ms_a, "ms_inversion": ms_i, "value": n / (ms_a + ms_i) / 1e3, "unit": UNIT,
                   "analysis_Msamples_s": n / ms_a / 1e3, "timing": "20 calls replayed as one CUDA graph (device time per call)",
                   "roofline": {"bound": "fp32 FMA (analysis)", "achieved": flop_a / (ms_a * 1e-3) / 1e12, "peak": 74.0,
                                "unit": "TFLOP/s", "frac": flop_a / (ms_a * 1e-3) / 1e12 / 74.0,
                                "peak_source": "nominal fp32 FMA roof (SURVEY 7 hard part 2)",
                                "hbm_frac_round_trip": by / ((ms_a + ms_i) * 1e-3) / 1e9 / peak}}
    if cpu:
        res["cfg1"]["cpu_baseline"] = _cpu_cfg_rate("cfg1", {"n": n, "taps": h, "taper": orc.window_tukey(256, 48)}, workers)
    del x, ch, y, ana, inv
    # ---- cfg3: SKA-Mid, 4096 ch, OS 8/7, 100 353-tap two-stage design, padded analysis, Nf=512 ov=128 ------
    hm = design_pfb_fir_filter_two_stage(4096, 8, 7)
    n = 1 << 24
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    xm = torch.view_as_complex(torch.randn((n, 2, 2), device=dev, generator=g))
    ana = ops.AnalysisPlan(4096, "8/7", hm, "polyphase_analysis_padded", device=dev)
    inv = ops.InversePlan(4096, "8/7", 512, 128, "tukey", deripple_taps=hm, device=dev)
    chm = ana(xm)
    ym = inv(chm)
    ms_a = _time_cuda(lambda: ana(xm, out=chm), 5)
    ms_i = _time_cuda(lambda: inv(chm, out=ym), 5)
    by = (xm.numel() + 2 * chm.numel() + ym.numel()) * cb
    res["cfg3"] = {"workload": f"SKA-Mid: 4096 ch, OS 8/7, {hm.size} taps (padded analysis), Nf=512 ov=128 tukey+deripple, "
                               f"2 pol x 2^24 complex64", "nblocks": int(chm.shape[0]),
                   "ms_analysis": ms_a, "ms_inversion": ms_i, "value": 2 * n / (ms_a + ms_i) / 1e3, "unit": UNIT,
                   "roofline": {"bound": "hbm", "achieved": by / ((ms_a + ms_i) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": by / ((ms_a + ms_i) * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes": by, "what": "round trip, channelized data materialised (SURVEY 8d)"}}
    if cpu:
        res["cfg3"]["cpu_baseline"] = _cpu_cfg_rate("cfg3", {"n": 1 << 22, "taps": hm, "taper": orc.window_tukey(512, 128)},
                                                    min(workers, 16))
    del xm, chm, ym, ana, inv
    torch.cuda.empty_cache()
    # ---- fp64 mode of the bench workload (north star: <= 1e-6 in an fp64 mode): generic kernels, complex128 -----
    n = 1 << 22
    x64 = torch.view_as_complex(torch.randn((n, 2, 2), device=dev, dtype=torch.float64, generator=g))
    ana = ops.AnalysisPlan(NCHAN, OS, h, dtype=torch.complex128, device=dev)
    inv = ops.InversePlan(NCHAN, OS, NFFT, OVERLAP, "tukey", deripple_taps=h, dtype=torch.complex128, device=dev)
    c64 = ana(x64)
    y64 = inv(c64)
    ms_a = _time_cuda(lambda: ana(x64, out=c64), 5)
    ms_i = _time_cuda(lambda: inv(c64, out=y64), 5)
    by = (x64.numel() + 2 * c64.numel() + y64.numel()) * 16
    res["cfg2_fp64_mode"] = {"workload": "the bench workload in complex128 (generic kernels), 2 pol x 2^22",
                             "ms_analysis": ms_a, "ms_inversion": ms_i, "value": 2 * n / (ms_a + ms_i) / 1e3, "unit": UNIT,
                             "roofline": {"bound": "hbm", "achieved": by / ((ms_a + ms_i) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": by / ((ms_a + ms_i) * 1e-3) / 1e9 / peak}}
    del x64, c64, y64, ana, inv
    torch.cuda.empty_cache()
    # ---- cfg5: batched purity sweep, impulse / tone / comb thirds, 147 456-sample vectors -------------------
    from ska_pst_dsp_model_b200.batch import BatchPurity
    vb = 128
    bp = BatchPurity(OS, NFFT, OVERLAP, "tukey", True, NCHAN, h, 3, device=dev, vectors_per_batch=vb)
    n_test = 11 * vb                                   # 3 x 1408 = 4224 vectors, whole chunks
    bp.reserve(n_test)                                 # pinned staging for the whole sweep (allocation is not the sweep)
    bp.sweep(vb, 3)                                    # warm-up: scratch sized, one graph per kind captured
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rep = bp.sweep(n_test, 3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nvec = sum(len(v) for v in rep.values())
    samples = nvec * bp.n_samples
    by = samples * (8 + 2 * 8 * 4.0 / 3.0 + 8)         # generate-in-place input, channelized write + read, output write
    res["cfg5"] = {"workload": f"purity sweep: {nvec} vectors (impulse, tone, comb thirds) of {bp.n_samples} samples, "
                               f"chunks of {vb} vectors, one analysis + one inversion pass per chunk, metrics on device",
                   "seconds": dt, "vectors_per_s": nvec / dt, "value": samples / dt / 1e6, "unit": UNIT,
                   "worst_max_spurious_dB": {"impulse": max(r["max_spurious_power"] for r in rep["test_time_domain_impulse"]
                                                            if bp.total_sample_shift < r["arg"] < bp.n_cmp),
                                             "tone": max(r["max_spurious_power"] for r in rep["test_complex_sinusoid"])},
                   "roofline": {"bound": "hbm", "achieved": by / dt / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": by / dt / 1e9 / peak,
                                "what": "wall clock of the whole sweep (host loop + graph replays + 16 doubles per vector D2H) "
                                        "against the round-trip bytes of the streaming kernels"}}
    if cpu:
        res["cfg5"]["cpu_baseline"] = _cpu_purity_rate(h, workers)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-samples", type=int, default=24, help="samples per pol per GPU (default 2^24)")
    ap.add_argument("--cpu-log2-samples", type=int, default=0,
                    help="CPU sample per pol (default: the whole per-GPU config for --impl reference, 2^22 for the cpu_baseline leg)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="length of the extra sustained run (0 = skip)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the cfg1 / cfg3 / cfg5 side measurements")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-segments", type=int, default=8, help="base segment count of the host round trip (ramped head and tail add a few)")
    ap.add_argument("--scratch-mib", type=int, default=0, help="override the inversion scratch cap (default 1 GiB: one chunk)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-int16", action="store_true", help="skip the int16-payload end-to-end variant")
    ap.add_argument("--no-fused", action="store_true", help="skip the fused analysis->inversion side measurement")
    ap.add_argument("--halo", default="auto", choices=["auto", "peer", "exchange"],
                    help="N > 1: read the neighbour's halo in place through a mapped peer pointer (default when available) "
                         "or exchange it with NCCL send/recv")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
