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


class TwoStageFilterBank:
    """Stage 1 on polarisation 0 only, one stage-2 FilterBank per coarse channel (they share one
    plan), optional critical chomp and single-channel output (TwoStageFilterBank.m:58-116)."""

    def __init__(self, cfg1: dict, cfg2: dict = None, critical=False, single=False):
        cfg2 = cfg2 or cfg1
        self.stage1 = FilterBank(**cfg1)
        first = FilterBank(**cfg2)
        self.stage2 = [first] + [FilterBank(**cfg2, plan=first.plan) for _ in range(self.stage1.n_chan - 1)]
        self.critical, self.single = critical, single

    def execute(self, x):
        torch = _torch()
        out1 = self.stage1.execute(x[:, 0:1].contiguous())            # :69 pol 1 only
        os_ = self.stage1.os_factor
        nch1 = self.stage1.n_chan
        nch2_orig = self.stage2[0].n_chan
        nch2 = nch2_orig * os_.de // os_.nu if self.critical else nch2_orig
        offset = nch2_orig - nch2
        if self.single:
            nch1 = 1
        out = None
        h = nch2 // 2
        for ich in range(nch1):
            tmp = self.stage2[ich].execute(out1[:, ich, 0:1].contiguous())
            if out is None:
                out = torch.zeros((tmp.shape[0], nch1 * nch2, 1), dtype=torch.complex64, device=tmp.device)   # :99 'single'
            out[:, ich * nch2: ich * nch2 + h] = tmp[:, :h]                                                    # :104
            out[:, ich * nch2 + h - 1: (ich + 1) * nch2] = tmp[:, h - 1 + offset: nch2 + offset]               # :105
        return out


class TwoStageInverseFilterBank:
    """Inverts stage 2 only, per output coarse channel (TwoStageInverseFilterBank.m:75-157)."""

    def __init__(self, cfg2: dict, nch2: int, stage2_nchan: int, combine=1, single=False):
        self.cfg2 = cfg2
        self.nch2, self.stage2_nchan = nch2, stage2_nchan
        self.combine, self.single = combine, single
        self.stage2 = {}

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
        out = None
        for ich in range(nch_out):
            if ich not in self.stage2:
                self.stage2[ich] = InverseFilterBank(**self.cfg2)
                if ich > 0:
                    self.stage2[ich]._plans = self.stage2[0]._plans          # plans are stateless: share them
            ifb = self.stage2[ich]
            ifb.critical, ifb.combine = critical, self.combine
            tmp = ifb.execute(x[:, ich * nch_in: (ich + 1) * nch_in, 0:1].contiguous())
            if out is None:
                out = torch.zeros((tmp.shape[0], nch_out, 1), dtype=tmp.dtype, device=tmp.device)
            out[:, ich, 0] = tmp[:, 0]
        return out
