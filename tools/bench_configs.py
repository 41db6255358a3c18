#!/usr/bin/env python
"""Timings of the other BASELINE configs (parity-test cases, not the bench line):
cfg1 (8 ch, 8/7, PST taps, 2^18 impulse), cfg3 (Mid: 4096 ch, 8/7, 100 353-tap two-stage design,
padded analysis + inversion Nf=512/ov=128), cfg5 (batched purity sweep)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from ska_pst_dsp_model_b200 import ops
from ska_pst_dsp_model_b200.batch import BatchPurity
from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter_two_stage, load_pst_taps


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
h = load_pst_taps(normalise=True)
# cfg1
n = 1 << 18
x = torch.zeros((n, 1), dtype=torch.complex64, device="cuda")
x[int(0.11 * n)] = 1
ana = ops.AnalysisPlan(8, "8/7", h)
inv = ops.InversePlan(8, "8/7", 256, 48, "tukey", deripple_taps=h)
ms_a = timeit(lambda: ana(x))
ch = ana(x)
ms_i = timeit(lambda: inv(ch))
res["cfg1"] = {"shape": "8 ch, OS 8/7, 3072 taps, 2^18 x 1 pol", "analysis_ms": ms_a, "inversion_ms": ms_i,
               "analysis_Msamples_s": n / ms_a / 1e3, "round_trip_Msamples_s": n / (ms_a + ms_i) / 1e3,
               "kernels": "generic (FMA-bound: 384 taps/arm)"}
# cfg3 (Mid)
if "--skip-mid" not in sys.argv:
    t0 = time.time()
    hm = design_pfb_fir_filter_two_stage(4096, 8, 7)
    t_design = time.time() - t0
    n = 1 << 22
    xm = torch.view_as_complex(torch.randn((n, 2, 2), device="cuda"))
    ana = ops.AnalysisPlan(4096, "8/7", hm, "polyphase_analysis_padded")
    t0 = time.time()
    inv = ops.InversePlan(4096, "8/7", 512, 128, "tukey", deripple_taps=hm)
    t_plan = time.time() - t0
    ms_a = timeit(lambda: ana(xm), 2)
    chm = ana(xm)
    ms_i = timeit(lambda: inv(chm), 2)
    res["cfg3"] = {"shape": f"4096 ch, OS 8/7, {hm.size} taps (padded), Nf=512, ov=128, 2^22 x 2 pol",
                   "filter_design_s": t_design, "inverse_plan_s": t_plan, "analysis_ms": ms_a, "inversion_ms": ms_i,
                   "analysis_Msamples_s": 2 * n / ms_a / 1e3, "round_trip_Msamples_s": 2 * n / (ms_a + ms_i) / 1e3,
                   "nblocks": int(chm.shape[0]), "kernels": "pfb_fir_phases + pfb_rowfft4096 (analysis), pfb_rowfft4096 + generic resample (inversion)"}
    del xm, chm, ana, inv
# two-stage: the whole chain of one block as ONE captured CUDA graph
# (filterbank.TwoStageGraph): stage 1 -> batched stage 2 over all coarse channels -> chomp -> batched stage-2 inversion.
from ska_pst_dsp_model_b200.filterbank import TwoStageGraph
n2 = 1 << 23
x2 = torch.view_as_complex(torch.randn((n2, 1, 2), device="cuda"))
cfg = dict(filt_coeff=h, n_chan=256, os_factor="4/3")
icfg = dict(filt_coeff=h, n_fft=128, overlap=16, os_factor="4/3", temporal_taper="tukey")
for critical, combine in ((False, 1), (True, 16)):
    tg = TwoStageGraph(cfg, cfg, icfg, n2, critical=critical, combine=combine)
    ms = timeit(lambda: tg.run(x2), 5)
    res[f"two_stage_low_graph_critical{int(critical)}_combine{combine}"] = {
        "shape": "2^23 x 1 pol -> 256 coarse x 256 fine channels (4/3 x 4/3, PST taps) -> stage-2 inversion (Nf=128, ov=16); "
                 "one CUDA graph replay per block (incl. the 64 MiB input copy)",
        "ms_per_block": ms, "Msamples_s": n2 / ms / 1e3, "fine_shape": list(tg.fine.shape), "out_shape": list(tg.y.shape)}
    del tg
del x2
if "--skip-mid" not in sys.argv and "--two-stage-mid" in sys.argv:
    # SKA-Mid two-stage analysis at scale (TwoStageFilterBank.m:58-116 with the mid config for both stages): 2^29 samples
    # -> 4096 coarse channels x 149 796 samples -> 41 stage-2 blocks of 4096 fine channels each.  The stage-2 INVERSION
    # needs >= 512 stage-2 blocks per coarse channel, i.e. 6.6e9 input samples (53 GB): not run (nor can sgcht.m's
    # 2 x 64 Mi-sample run produce it, sgcht.m:481-484).
    hm2 = design_pfb_fir_filter_two_stage(4096, 8, 7)
    cfgm = dict(filt_coeff=hm2, n_chan=4096, os_factor="8/7", analysis_function="polyphase_analysis_padded")
    nm = 1 << 29
    xm2 = torch.view_as_complex(torch.randn((nm, 1, 2), device="cuda"))
    tgm = TwoStageGraph(cfgm, cfgm, None, nm, critical=True, invert=False)
    ms = timeit(lambda: tgm.run(xm2), 2)
    res["two_stage_mid_graph_analysis"] = {"shape": "2^29 x 1 pol -> 4096 coarse x 3584 kept fine channels (critical), one graph",
                                           "ms_per_block": ms, "Msamples_s": nm / ms / 1e3, "fine_shape": list(tgm.fine.shape)}
    del xm2, tgm
# file -> file through the streaming operators: 2^24 x 2 pol complex64 DADA file (256 MiB) channelized
# to 341 MiB and inverted to 255 MiB, pieces read into pinned memory / written by I/O threads
import tempfile
from ska_pst_dsp_model_b200.file_streaming import stream_channelize, stream_synthesize
from ska_pst_dsp_model_b200.format_handler import PSRFormatChannelizer, PSRFormatSynthesizer
from ska_pst_dsp_model_b200.formats import DADAFile
from ska_pst_dsp_model_b200.fft_windows import tukey_window
with tempfile.TemporaryDirectory() as td:
    nf = 1 << 24
    xf = (np.random.default_rng(0).standard_normal((nf, 1, 2, 2)).astype(np.float32)).view(np.complex64)[..., 0]
    DADAFile(os.path.join(td, "in.dump"), data=xf).dump_data()
    chz = PSRFormatChannelizer("4/3", 256, h)
    syn = PSRFormatSynthesizer(48, tukey_window(256, 48), 256, apply_deripple=True)
    stream_channelize(chz, os.path.join(td, "in.dump"), td, "w.dump", blocks_per_piece=4096)      # warm-up (plans, page cache)
    t0 = time.time()
    ch_f = stream_channelize(chz, os.path.join(td, "in.dump"), td, "ch.dump", blocks_per_piece=4096)
    t1 = time.time()
    stream_synthesize(syn, ch_f.file_path, td, "syn.dump", blocks_per_piece=32)
    t2 = time.time()
    res["file_streaming"] = {"shape": "2^24 x 2 pol complex64 DADA file -> channelized file -> inverted file (tmpfs/disk of the box)",
                             "channelize_s": t1 - t0, "synthesize_s": t2 - t1,
                             "channelize_Msamples_s": nf / (t1 - t0) / 1e6, "synthesize_Msamples_s": nf / (t2 - t1) / 1e6}
    del xf
# cfg5: batched purity sweep
nvec = int(os.environ.get("NVEC", "512"))
bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=32)
bp.sweep(64, 3)          # warm-up with full chunks: graphs for the chunk size are captured here
torch.cuda.synchronize()
t0 = time.time()
rep = bp.sweep(nvec, 3)
torch.cuda.synchronize()
dt = time.time() - t0
worst_t = max(r["max_spurious_power"] for r in rep["test_time_domain_impulse"][:-1])
worst_f = max(r["max_spurious_power"] for r in rep["test_complex_sinusoid"])
res["cfg5"] = {"shape": f"{2 * nvec} vectors of 147456 samples (impulse + tone sweep, purity.py grid)",
               "seconds": dt, "vectors_per_s": 2 * nvec / dt, "Msamples_s": 2 * nvec * 147456 / dt / 1e6,
               "worst_max_spurious_dB_time": worst_t, "worst_max_spurious_dB_freq": worst_f}
print(json.dumps(res, indent=1))
