"""Streaming operator classes on the GPU, mirroring the Matlab classes of the reference:

* ``FilterBank``                 <-> matlab/FilterBank.m:65-128
* ``InverseFilterBank``          <-> matlab/InverseFilterBank.m:63-137
* ``TwoStageFilterBank``         <-> matlab/TwoStageFilterBank.m:58-116
* ``TwoStageInverseFilterBank``  <-> matlab/TwoStageInverseFilterBank.m:75-157

Data are torch CUDA tensors in TFP order: time series ``[ndat, npol]``, channelized
``[ndat, nchan, npol]`` (the Matlab ``(npol, nchan, ndat)`` arrays are the same bytes).
The tail carry between ``execute`` calls is kept on the device.
"""
from __future__ import annotations

import numpy as np

from . import ops
from .fft_windows import make_window
from .lib import PfbError
from .rational import Rational


def _torch():
    import torch
    return torch


def _matlab_round(x):
    """Matlab round(): half away from zero, applied to real and imaginary parts."""
    torch = _torch()
    xr = torch.view_as_real(x)
    return torch.view_as_complex(torch.sign(xr) * torch.floor(torch.abs(xr) + 0.5))


def _std_all(x):
    """sqrt(var(x, 0, 'all')) of Matlab for complex data (N-1 normalisation)."""
    torch = _torch()
    n = x.numel()
    mean = x.mean()
    return torch.sqrt(((x - mean).abs() ** 2).sum() / (n - 1))


class FilterBank:
    def __init__(self, filt_coeff, n_chan, os_factor, analysis_function="polyphase_analysis", rnd_input=False,
                 rms_input=0.0, rnd_output=False, rms_output=0.0, dtype=None, device=None, plan=None):
        torch = _torch()
        self.filt_coeff = np.asarray(filt_coeff, dtype=np.float64).ravel()
        self.n_chan = int(n_chan)
        self.os_factor = Rational.from_str(os_factor)
        self.analysis_function = analysis_function
        self.rnd_input, self.rms_input = rnd_input, rms_input
        self.rnd_output, self.rms_output = rnd_output, rms_output
        self.dtype = dtype or torch.complex64
        self.plan = plan or ops.AnalysisPlan(self.n_chan, self.os_factor, self.filt_coeff, analysis_function,
                                             dtype=self.dtype, device=device)
        self.input_buffer = None
        self.buffered_samples = 0
        self._first_call = True

    def execute(self, x):
        torch = _torch()
        if self.rnd_input:                                   # FilterBank.m:75-83
            scale = 1.0
            if self.rms_input > 0:
                scale = self.rms_input / _std_all(x)
            x = _matlab_round(scale * x)
        if self.buffered_samples > 0:                        # :85-88
            x = torch.cat([self.input_buffer, x], dim=0)
            self.buffered_samples = 0
        x = x.contiguous()
        n_in = x.shape[0]
        out = self.plan(x, first_call=self._first_call)      # :91
        self._first_call = False
        rem = out.shape[0] % self.os_factor.nu               # :93-104
        if rem:
            out = out[: out.shape[0] - rem]
        if self.rnd_output:                                  # :106-113
            scale = 1.0
            if self.rms_output:
                scale = self.rms_output / _std_all(out)
            out = _matlab_round(scale * out)
        input_idat = out.shape[0] * self.n_chan * self.os_factor.de // self.os_factor.nu   # :119
        self.buffered_samples = n_in - input_idat
        if self.buffered_samples > 0:
            self.input_buffer = x[input_idat:].clone()
        return out


class InverseFilterBank:
    def __init__(self, filt_coeff, n_fft, overlap, os_factor, temporal_taper="tukey", critical=False, combine=1,
                 sample_offset=0, spectral_taper=None, dtype=None, device=None):
        torch = _torch()
        if spectral_taper not in (None, "no_window", "identity", "hann"):
            raise PfbError(f"unknown spectral taper {spectral_taper!r}")
        self.spectral_taper = spectral_taper if spectral_taper == "hann" else None
        self.filt_coeff = filt_coeff
        self.n_fft, self.overlap = int(n_fft), int(overlap)
        self.os_factor = Rational.from_str(os_factor)
        self.temporal_taper = make_window(temporal_taper or "no_window", self.n_fft, self.overlap)
        self.critical, self.combine, self.sample_offset = critical, combine, sample_offset
        self.dtype = dtype or torch.complex64
        self.device = device
        self._plans = {}
        self.input_buffer = None
        self.buffered_samples = 0

    def _plan(self, nchan):
        key = (nchan, self.critical, self.combine)
        if key not in self._plans:
            # InverseFilterBank.m:90 forces deripple off; spans_nyquist = ~critical (:89)
            self._plans[key] = ops.InversePlan(nchan, self.os_factor, self.n_fft, self.overlap, self.temporal_taper,
                                               deripple_taps=None, spans_nyquist=not self.critical,
                                               combine=self.combine, dtype=self.dtype, device=self.device)
            if self.spectral_taper:
                self._plans[key].set_spectral_taper(self.spectral_taper)
        return self._plans[key]

    def execute(self, x):
        torch = _torch()
        if self.buffered_samples > 0:
            x = torch.cat([self.input_buffer, x], dim=0)
            self.buffered_samples = 0
        if not torch.is_complex(x):
            raise ValueError("polyphase_synthesis input data are real-valued!")      # InverseFilterBank.m:83-85
        x = x.contiguous()
        n_dat, n_chan, _ = x.shape
        out = self._plan(n_chan)(x, sample_offset=self.sample_offset + 1)
        nu, de = self.os_factor.nu, self.os_factor.de
        while True:                                                                   # :104-124
            input_idat = out.shape[0] * nu // (n_chan * de)
            self.buffered_samples = n_dat - input_idat
            rem = self.buffered_samples % nu
            if rem == 0:
                break
            self.buffered_samples += nu - rem
            input_idat = n_dat - self.buffered_samples
            out = out[: input_idat * (n_chan * de) // nu]
        if self.buffered_samples > 0:
            self.input_buffer = x[input_idat:].clone()
        return out


class _BatchedFilterBank(FilterBank):
    """FilterBank.m semantics applied to many single-pol streams at once (one tail buffer for all)."""

    def execute_batched(self, x):
        torch = _torch()
        if self.rnd_input:
            scale = 1.0
            if self.rms_input > 0:
                scale = self.rms_input / _std_all(x)
            x = _matlab_round(scale * x)
        if self.buffered_samples > 0:
            x = torch.cat([self.input_buffer, x], dim=0)
            self.buffered_samples = 0
        x = x.contiguous()
        n_in = x.shape[0]
        out = self.plan.batched(x, first_call=self._first_call, coarse_major=True)     # [k2][stream][chan]
        self._first_call = False
        rem = out.shape[0] % self.os_factor.nu
        if rem:
            out = out[: out.shape[0] - rem]
        if self.rnd_output:
            # FilterBank.m:106-113 scales each stage-2 filterbank by its own output rms
            if self.rms_output:
                n = out.shape[0] * out.shape[2]
                mean = out.mean(dim=(0, 2), keepdim=True)
                std = torch.sqrt(((out - mean).abs() ** 2).sum(dim=(0, 2), keepdim=True) / (n - 1))
                out = out * (self.rms_output / std)
            out = _matlab_round(out)
        input_idat = out.shape[0] * self.n_chan * self.os_factor.de // self.os_factor.nu
        self.buffered_samples = n_in - input_idat
        if self.buffered_samples > 0:
            self.input_buffer = x[input_idat:].clone()
        return out


class TwoStageFilterBank:
    """Stage 1 on polarisation 0 only, then one stage-2 filterbank per coarse channel - here a single
    batched launch over all coarse channels - optional critical chomp and single-channel output
    (TwoStageFilterBank.m:58-116)."""

    def __init__(self, cfg1: dict, cfg2: dict = None, critical=False, single=False):
        cfg2 = cfg2 or cfg1
        self.stage1 = FilterBank(**cfg1)
        self.stage2 = _BatchedFilterBank(**cfg2)
        self.critical, self.single = critical, single

    def execute(self, x):
        torch = _torch()
        out1 = self.stage1.execute(x[:, 0:1].contiguous())            # :69 pol 1 only -> [k1][nch1][1]
        os_ = self.stage1.os_factor
        nch1 = 1 if self.single else self.stage1.n_chan
        nch2_orig = self.stage2.n_chan
        nch2 = nch2_orig * os_.de // os_.nu if self.critical else nch2_orig
        offset = nch2_orig - nch2
        tmp = self.stage2.execute_batched(out1[:, :nch1, 0].contiguous())   # [k2][nch1][nch2_orig]
        h = nch2 // 2
        if offset:
            # :104-105 - keep [0, h) and [h-1+offset, nch2+offset); the second assignment overwrites index h-1
            idx = torch.cat([torch.arange(0, h - 1, device=tmp.device),
                             torch.arange(h - 1 + offset, nch2 + offset, device=tmp.device)])
            tmp = tmp.index_select(2, idx)
        out = tmp.reshape(tmp.shape[0], nch1 * nch2, 1).to(torch.complex64)   # :99 'single'
        return out.contiguous()


class TwoStageInverseFilterBank:
    """Inverts stage 2 only, one inversion per output coarse channel - here one batched launch
    (TwoStageInverseFilterBank.m:75-157)."""

    def __init__(self, cfg2: dict, nch2: int, stage2_nchan: int, combine=1, single=False):
        self.cfg2 = cfg2
        self.nch2, self.stage2_nchan = nch2, stage2_nchan
        self.combine, self.single = combine, single
        self.inv = None
        self.input_buffer = None
        self.buffered_samples = 0

    def execute(self, x):
        torch = _torch()
        os_ = Rational.from_str(self.cfg2["os_factor"])
        nchan = x.shape[1]
        nch_out = nchan // self.nch2
        if self.nch2 == os_.normalize(self.stage2_nchan):
            critical = True
        elif self.nch2 == self.stage2_nchan:
            critical = False
            if self.combine > 1:
                raise ValueError("TwoStageInverseFilterBank::execute cannot combine oversampled coarse channels")
        else:
            raise ValueError("TwoStageInverseFilterBank::execute invalid nchan")
        nch_in = self.nch2 * self.combine
        nch_out //= self.combine
        if self.single:
            nch_out = 1
        if self.inv is None:
            self.inv = InverseFilterBank(**self.cfg2)
        ifb = self.inv
        ifb.critical, ifb.combine = critical, self.combine
        if self.buffered_samples > 0:
            x = torch.cat([self.input_buffer, x], dim=0)
            self.buffered_samples = 0
        x = x.contiguous()
        n_dat = x.shape[0]
        xg = x[:, : nch_out * nch_in, 0:1].reshape(n_dat, nch_out, nch_in, 1).contiguous()
        out = ifb._plan(nch_in).batched(xg, sample_offset=ifb.sample_offset + 1)        # [n_out][nch_out][1]
        nu, de = os_.nu, os_.de
        while True:                                   # InverseFilterBank.m:104-124 (same for every group)
            input_idat = out.shape[0] * nu // (nch_in * de)
            self.buffered_samples = n_dat - input_idat
            rem = self.buffered_samples % nu
            if rem == 0:
                break
            self.buffered_samples += nu - rem
            input_idat = n_dat - self.buffered_samples
            out = out[: input_idat * (nch_in * de) // nu]
        if self.buffered_samples > 0:
            self.input_buffer = x[input_idat:].clone()
        return out.contiguous()


class TwoStageGraph:
    """One block of the two-stage chain as a single captured launch graph: stage-1 analysis ->
    stage-2 analysis of all coarse channels (one batched launch) -> critical chomp -> stage-2 inversion of all coarse
    channels (one batched launch, optional ``combine``), i.e. ``TwoStageInverseFilterBank.execute(
    TwoStageFilterBank.execute(x))`` (TwoStageFilterBank.m:58-116, TwoStageInverseFilterBank.m:75-157) for the FIRST
    block of a stream of fixed length ``n_dat`` (no carried tails inside the graph: a streaming caller feeds blocks
    that already contain their history, like ``file_streaming``).  Buffers are static; ``run(x)`` copies the block in,
    replays the graph and returns views of the static outputs."""

    def __init__(self, cfg1: dict, cfg2: dict, icfg: dict, n_dat: int, critical=False, combine=1, invert=True):
        torch = _torch()
        self.fb = TwoStageFilterBank(cfg1, cfg2, critical=critical)
        nch2 = self.fb.stage2.n_chan
        os2 = self.fb.stage2.os_factor
        self.inv = TwoStageInverseFilterBank(icfg, nch2=(nch2 * os2.de // os2.nu if critical else nch2),
                                             stage2_nchan=nch2, combine=combine) if invert else None
        self.x = torch.zeros((int(n_dat), 1), dtype=torch.complex64, device="cuda")
        self._fresh()
        self.fine = self.fb.execute(self.x)                      # warm-up: plans built, scratch sized, shapes fixed
        self.y = self.inv.execute(self.fine) if self.inv is not None else None
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(self.graph, stream=cap, capture_error_mode="thread_local"):
            self._fresh()
            self.fine = self.fb.execute(self.x)
            self.y = self.inv.execute(self.fine) if self.inv is not None else None
        torch.cuda.current_stream().wait_stream(cap)

    def _fresh(self):
        for f in (self.fb.stage1, self.fb.stage2):
            f.buffered_samples, f.input_buffer, f._first_call = 0, None, True
        if self.inv is not None:
            self.inv.buffered_samples, self.inv.input_buffer = 0, None

    def run(self, x):
        self.x.copy_(x.reshape(self.x.shape))
        self.graph.replay()
        return self.fine, self.y
