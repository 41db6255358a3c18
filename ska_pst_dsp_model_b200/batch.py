"""Batched purity sweep on the device (BASELINE configs[4]).

The reference's purity driver (python/verify/purity.py:165-207) runs one
generate -> channelize -> synthesize -> compare pass per test vector through files on disk.
Here every vector is synthesised on the device (``pfb_generate_test_vector``), channelized and
inverted by the CUDA kernels, and reduced to the report's metrics by ``pfb_purity_metrics``;
only ~8 doubles per vector leave the GPU.  Two independent vectors ride in the two polarisation slots
of one [n][2] stream (``pfb_generate_test_vector_strided``), so every pair goes through the dual-pol
kernels in a single analysis + inversion pass.  Vectors are independent: across GPUs they are dealt
round-robin (replicas, no data-path collective) and the per-vector records are gathered at the end.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import lib as _l
from . import ops
from .fft_windows import make_window
from .rational import Rational

_REC = np.dtype([("a", "<u8"), ("b", "<u8"), ("n", "<i8"), ("stride_a", "<i8"), ("stride_b", "<i8"),
                 ("scale_b", "<f8"), ("out", "<u8")])


def _db(p):
    return 10.0 * np.log10(np.abs(p) + 1e-13)          # verify/util.py:39-43


class BatchPurity:
    def __init__(self, os_factor, input_fft_length, input_overlap, fft_window, deripple, channels, fir_taps,
                 blocks, analysis_function="polyphase_analysis", device=None, vectors_per_batch=16):
        import torch
        assert _REC.itemsize == 56
        self.torch = torch
        self.osf = Rational.from_str(os_factor)
        taps = np.asarray(fir_taps, dtype=np.float64).ravel()
        self.block_size = self.osf.normalize(input_fft_length) * channels          # purity.py:82
        self.fft_size = 2 * self.block_size
        self.n_samples = self.block_size * blocks
        self.total_sample_shift = self.osf.normalize(input_overlap) * channels + (taps.size - 1) // 2   # :86-88
        self.analysis = ops.AnalysisPlan(channels, self.osf, taps, analysis_function, device=device)
        self.inverse = ops.InversePlan(channels, self.osf, input_fft_length, input_overlap,
                                       make_window(fft_window, input_fft_length, input_overlap),
                                       deripple_taps=taps if deripple else None, device=device)
        self.device = self.analysis.device
        self.vb = vectors_per_batch
        dev, n = self.device, self.n_samples
        nb = self.analysis.nblocks(n)
        self.n_out = self.inverse.out_len(nb)
        self.vb += self.vb & 1                                    # whole pairs
        self.x = torch.empty((self.vb // 2, n, 2), dtype=torch.complex64, device=dev)
        self.ch = torch.empty((nb, channels, 2), dtype=torch.complex64, device=dev)
        self.y = torch.empty((self.vb // 2, self.n_out, 2), dtype=torch.complex64, device=dev)
        self.met = torch.zeros((self.vb, 2, 8), dtype=torch.float64, device=dev)
        self.n_cmp = min(n - self.total_sample_shift, self.n_out)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _metrics(self, records):
        torch = self.torch
        rec = torch.from_numpy(records.view(np.uint8).reshape(-1)).to(self.device)
        _l.check(_l.lib.pfb_purity_metrics(rec.data_ptr(), len(records), _l.PFB_DTYPE_C64, self._stream()))
        return rec      # keep alive until the stream has consumed it

    def _run(self, kind, params):
        torch = self.torch
        out = []
        esz = 8
        with torch.cuda.device(self.device):
            for b0 in range(0, len(params), self.vb):
                chunk = params[b0: b0 + self.vb]
                npair = (len(chunk) + 1) // 2
                for i, prm in enumerate(chunk):
                    # vector i -> polarisation slot i & 1 of pair i >> 1
                    _l.check(_l.lib.pfb_generate_test_vector_strided(kind, self.x[i >> 1].data_ptr() + (i & 1) * esz,
                                                                     self.n_samples, 2, _l.PFB_DTYPE_C64, *prm,
                                                                     self._stream()))
                if len(chunk) & 1:                    # odd tail: the spare slot repeats the last vector
                    _l.check(_l.lib.pfb_generate_test_vector_strided(kind, self.x[npair - 1].data_ptr() + esz,
                                                                     self.n_samples, 2, _l.PFB_DTYPE_C64, *chunk[-1],
                                                                     self._stream()))
                for pi in range(npair):
                    self.analysis(self.x[pi], out=self.ch)
                    self.inverse(self.ch, out=self.y[pi])
                shift = self.total_sample_shift
                recs = np.zeros(2 * len(chunk), dtype=_REC)
                keep = []

                def xa(i):
                    return self.x[i >> 1].data_ptr() + (shift * 2 + (i & 1)) * esz

                def yb(i):
                    return self.y[i >> 1].data_ptr() + (i & 1) * esz

                if kind == 0:
                    for i in range(len(chunk)):
                        recs[2 * i] = (xa(i), yb(i), self.n_cmp, 2, 2, 1.0, self.met[i, 0].data_ptr())
                        recs[2 * i + 1] = recs[2 * i]
                        recs[2 * i + 1]["out"] = self.met[i, 1].data_ptr()
                else:
                    fa = torch.fft.fft(self.x[:npair, shift: shift + self.fft_size] / self.fft_size, dim=1).contiguous()
                    fb = torch.fft.fft(self.y[:npair, : self.fft_size] / self.fft_size, dim=1).contiguous()
                    keep = [fa, fb]
                    for i in range(len(chunk)):
                        recs[2 * i] = (xa(i), yb(i), self.n_cmp, 2, 2, 1.0, self.met[i, 0].data_ptr())
                        recs[2 * i + 1] = (fa[i >> 1].data_ptr() + (i & 1) * esz, fb[i >> 1].data_ptr() + (i & 1) * esz,
                                           self.fft_size, 2, 2, 1.0, self.met[i, 1].data_ptr())
                keep.append(self._metrics(recs))
                m = self.met[: len(chunk)].cpu().numpy()
                del keep
                for i, prm in enumerate(chunk):
                    t, s = m[i, 0], m[i, 1]
                    n_s = s[7]
                    out.append({
                        "mean_diff": float(t[0] / t[7]), "total_diff": float(t[0]),
                        "max_spurious_power": float(_db(s[4])),
                        "total_spurious_power": float(_db(s[2] - s[3])),
                        "mean_spurious_power": float(_db((s[2] - s[3]) / n_s)),
                        "arg": prm[0],
                    })
        return out

    def temporal(self, offsets, width=1):
        """purity.py:209-234 (test_time_domain_impulse)."""
        return self._run(0, [(float(int(o)), float(width), 0.0) for o in offsets])

    def spectral(self, frequencies, phase=np.pi / 4, bin_offset=0.0):
        """purity.py:236-263 (test_complex_sinusoid)."""
        return self._run(1, [(float(f), float(phase), float(bin_offset)) for f in frequencies])

    def sweep(self, n_test, blocks, rank=0, world=1):
        """The reference's parameter grid (purity.py:94-99), this rank's round-robin share."""
        offsets = np.linspace(1, self.n_samples, n_test).astype(int)[rank::world]
        freqs = (np.linspace(1, self.block_size, n_test).astype(int) * blocks)[rank::world]
        return {"test_time_domain_impulse": self.temporal(offsets), "test_complex_sinusoid": self.spectral(freqs)}


def gather_reports(report, world):
    """Merge the per-rank reports (all_gather_object; a few scalars per vector)."""
    if world == 1:
        return report
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, report)
    merged = {}
    for key in report:
        rows = [r for p in parts for r in p[key]]
        merged[key] = sorted(rows, key=lambda r: r["arg"])
    return merged
