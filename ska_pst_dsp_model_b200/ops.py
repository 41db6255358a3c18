"""Host-side operators over the C ABI: device buffers are torch tensors, the
arithmetic is the CUDA library's.  Mirrors the reference operator signatures:

* ``AnalysisPlan``  <->  ``polyphase_analysis[_padded|_lowcbf](in, filt, block, os_factor)``
  (matlab/polyphase_analysis.m:1-7) / ``FilterBank.pfb_analysis`` slot (FilterBank.m:6,36,91)
* ``InversePlan``   <->  ``polyphase_synthesis(in, spans, Nf, os, deripple, sample_offset,
  input_overlap, temporal_taper, spectral_taper, combine)`` (polyphase_synthesis.m:1-13)

Data layout is DADA TFP: analysis in ``[ndat, npol]`` -> out ``[nblocks, nchan, npol]``;
inversion in ``[ndat, nchan, npol]`` -> out ``[nout, npol]``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import lib as _l
from .fft_windows import make_window
from .rational import Rational

_DT = {torch.complex64: _l.PFB_DTYPE_C64, torch.complex128: _l.PFB_DTYPE_C128}


def _dptr(arr):
    if arr is None:
        return None
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).ravel())
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_tensor(x, dtype, device, ndim):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("expected a CUDA tensor (there is no CPU path)")
    if x.dtype != dtype:
        raise TypeError(f"expected dtype {dtype}, got {x.dtype}")
    if x.device != device:
        raise ValueError(f"tensor on {x.device}, plan on {device}")
    if x.dim() != ndim or not x.is_contiguous():
        raise ValueError(f"expected a contiguous {ndim}-d tensor, got shape {tuple(x.shape)}")


class AnalysisPlan:
    """Oversampled PFB analysis on one GPU."""

    def __init__(self, nchan, os_factor, taps, variant="polyphase_analysis",
                 dtype=torch.complex64, device=None):
        self.os_factor = Rational.from_str(os_factor)
        self.nchan = int(nchan)
        self.variant = _l.VARIANTS[variant] if isinstance(variant, str) else int(variant)
        self.dtype = dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.taps = np.asarray(taps, dtype=np.float64).ravel()
        keep, ptr = _dptr(self.taps)
        self._h = ctypes.c_void_p()
        _l.check(_l.lib.pfb_analysis_plan_create(ctypes.byref(self._h), self.variant, self.nchan,
                                                 self.os_factor.nu, self.os_factor.de, ptr, self.taps.size,
                                                 _DT[dtype], self.device.index))
        self.nchan_out = _l.lib.pfb_analysis_nchan_out(self._h)
        self.step = self.nchan * self.os_factor.de // self.os_factor.nu
        self.fl = -(-self.taps.size // self.nchan) * self.nchan

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _l is not None and getattr(_l, "lib", None) is not None:     # (module may be gone at interpreter exit)
            _l.lib.pfb_analysis_plan_destroy(h)

    def nblocks(self, n_dat, first_call=True):
        return int(_l.lib.pfb_analysis_nblocks(self._h, int(n_dat), int(first_call)))

    def __call__(self, x, first_call=True, out=None):
        _check_tensor(x, self.dtype, self.device, 2)
        n_dat, npol = x.shape
        nb = self.nblocks(n_dat, first_call)
        if out is None:
            out = torch.empty((nb, self.nchan_out, npol), dtype=self.dtype, device=self.device)
        else:
            _check_tensor(out, self.dtype, self.device, 3)
            if tuple(out.shape) != (nb, self.nchan_out, npol):
                raise ValueError(f"out must have shape {(nb, self.nchan_out, npol)}")
        with torch.cuda.device(self.device):
            _l.check(_l.lib.pfb_analysis_execute(self._h, x.data_ptr(), out.data_ptr(), npol, n_dat,
                                                 int(first_call), _stream_ptr(self.device)))
        return out

    def batched(self, x, first_call=True, coarse_major=True):
        """``x``: [n_dat, nstreams] independent single-pol streams (e.g. the coarse channels of a first
        stage).  Returns [nblocks, nstreams, nchan_out] (coarse_major) or [nblocks, nchan_out, nstreams]."""
        _check_tensor(x, self.dtype, self.device, 2)
        n_dat, nstreams = x.shape
        nb = self.nblocks(n_dat, first_call)
        shape = (nb, nstreams, self.nchan_out) if coarse_major else (nb, self.nchan_out, nstreams)
        out = torch.empty(shape, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _l.check(_l.lib.pfb_analysis_execute_batched(self._h, x.data_ptr(), out.data_ptr(), nstreams, n_dat,
                                                         int(first_call), 1 if coarse_major else 0,
                                                         _stream_ptr(self.device)))
        return out

    def segment(self, x_local, sample0, k_first, nblk, out=None, first_call=True):
        """Blocks [k_first, k_first+nblk) of the stream whose sample ``sample0`` is x_local[0]."""
        _check_tensor(x_local, self.dtype, self.device, 2)
        n_local, npol = x_local.shape
        if out is None:
            out = torch.empty((nblk, self.nchan_out, npol), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _l.check(_l.lib.pfb_analysis_execute_segment(self._h, x_local.data_ptr(), n_local, int(sample0),
                                                         out.data_ptr(), npol, int(k_first), int(nblk),
                                                         int(first_call), _stream_ptr(self.device)))
        return out


    def segment_peer(self, x_own, next_ptr, n_next, sample0, k_first, nblk, out=None, first_call=True):
        """``segment`` with the stream in two buffers: ``x_own`` and the ``n_next`` samples that follow it at device
        address ``next_ptr`` (a peer GPU's buffer mapped into this process): the halo is read in place over NVLink."""
        _check_tensor(x_own, self.dtype, self.device, 2)
        n_own, npol = x_own.shape
        if out is None:
            out = torch.empty((nblk, self.nchan_out, npol), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _l.check(_l.lib.pfb_analysis_execute_segment_peer(self._h, x_own.data_ptr(), n_own, ctypes.c_void_p(int(next_ptr)),
                                                              int(n_next), int(sample0), out.data_ptr(), npol, int(k_first),
                                                              int(nblk), int(first_call), _stream_ptr(self.device)))
        return out


class InversePlan:
    """Golden FFT-based PFB inversion on one GPU."""

    def __init__(self, nchan, os_factor, input_fft_length, input_overlap, temporal_taper="no_window",
                 deripple_taps=None, spans_nyquist=True, combine=1, dtype=torch.complex64, device=None):
        self.os_factor = Rational.from_str(os_factor)
        self.nchan, self.nfft, self.overlap = int(nchan), int(input_fft_length), int(input_overlap)
        self.dtype = dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if isinstance(temporal_taper, str) or temporal_taper is None:
            temporal_taper = make_window(temporal_taper or "no_window", self.nfft, self.overlap)
        taper = np.asarray(temporal_taper, dtype=np.float64).ravel()
        if taper.size != self.nfft:
            raise ValueError("temporal taper must have input_fft_length points")
        _, tptr = k1 = _dptr(taper)
        if deripple_taps is not None:
            k2 = _dptr(deripple_taps)
            dptr, nd = k2[1], k2[0].size
        else:
            dptr, nd = None, 0
        self._h = ctypes.c_void_p()
        _l.check(_l.lib.pfb_inverse_plan_create(ctypes.byref(self._h), self.nchan, self.os_factor.nu,
                                                self.os_factor.de, self.nfft, self.overlap, tptr, dptr, nd,
                                                int(bool(spans_nyquist)), int(combine), _DT[dtype],
                                                self.device.index))
        self.keep = self.nfft - 2 * self.overlap
        self.fn_width = self.nfft * self.os_factor.de // self.os_factor.nu
        self.output_keep = (self.fn_width - 2 * (self.overlap * self.os_factor.de // self.os_factor.nu)) * self.nchan

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _l is not None and getattr(_l, "lib", None) is not None:
            _l.lib.pfb_inverse_plan_destroy(h)

    def nblocks(self, n_dat):
        return int(_l.lib.pfb_inverse_nblocks(self._h, int(n_dat)))

    def out_len(self, n_dat):
        return int(_l.lib.pfb_inverse_out_len(self._h, int(n_dat)))

    def set_scratch_bytes(self, nbytes):
        _l.check(_l.lib.pfb_inverse_set_scratch_bytes(self._h, int(nbytes)))

    def reserve(self, n_dat, npol=2, ngroups=1):
        """Size the plan's scratch for inputs of up to n_dat fine-channel samples before capturing a CUDA graph
        (a plan is single-stream, not thread-safe, and regrows its scratch with cudaFree + cudaMalloc)."""
        _l.check(_l.lib.pfb_inverse_reserve(self._h, int(npol), int(n_dat), int(ngroups)))

    def __call__(self, x, out=None, sample_offset=1):
        """sample_offset is 1-based like polyphase_synthesis.m:99."""
        _check_tensor(x, self.dtype, self.device, 3)
        if sample_offset > 1:
            x = x[sample_offset - 1:]
        n_dat, nchan, npol = x.shape
        if nchan != self.nchan:
            raise ValueError(f"plan built for {self.nchan} channels, data has {nchan}")
        n_out = self.out_len(n_dat)
        if out is None:
            out = torch.empty((n_out, npol), dtype=self.dtype, device=self.device)
        else:
            _check_tensor(out, self.dtype, self.device, 2)
            if tuple(out.shape) != (n_out, npol):
                raise ValueError(f"out must have shape {(n_out, npol)}")
        with torch.cuda.device(self.device):
            _l.check(_l.lib.pfb_inverse_execute(self._h, x.data_ptr(), out.data_ptr(), npol, n_dat,
                                                _stream_ptr(self.device)))
        return out

    def set_spectral_taper(self, taper):
        """Real window of nchan*FN_width points on the stitched spectrum (polyphase_synthesis.m:282);
        ``"hann"`` builds circshift(hann(len), len/2) as PFBWindow.m:79-94 does; None clears it."""
        if taper is None:
            _l.check(_l.lib.pfb_inverse_set_spectral_taper(self._h, None, 0))
            return
        n = self.nchan * self.fn_width
        if isinstance(taper, str):
            if taper != "hann":
                raise KeyError(f"unknown spectral taper {taper!r}")
            from .fft_windows import hann_window
            taper = np.roll(hann_window(n), n // 2)
        keep, ptr = _dptr(taper)
        _l.check(_l.lib.pfb_inverse_set_spectral_taper(self._h, ptr, keep.size))

    def set_profiling(self, on=True):
        _l.check(_l.lib.pfb_inverse_set_profiling(self._h, int(on)))

    def get_profile(self):
        """{'chan_idft': (ms, launches), 'resample': (ms, launches)} since the last call."""
        ms = (ctypes.c_double * 2)()
        n = (ctypes.c_longlong * 2)()
        _l.check(_l.lib.pfb_inverse_get_profile(self._h, ms, n))
        return {"chan_idft": (ms[0], n[0]), "resample": (ms[1], n[1])}

    def batched(self, x, sample_offset=1):
        """``x``: [n_dat, ngroups, nchan, npol] - ngroups independent channel sets (the coarse channels of a
        two-stage filterbank).  Returns the inverted series [n_out, ngroups, npol]."""
        _check_tensor(x, self.dtype, self.device, 4)
        if sample_offset > 1:
            x = x[sample_offset - 1:]
        n_dat, G, nchan, npol = x.shape
        if nchan != self.nchan:
            raise ValueError(f"plan built for {self.nchan} channels, data has {nchan}")
        nblk = self.nblocks(n_dat)
        keep2 = self.output_keep // self.nchan
        tmp = torch.empty((nblk * keep2, G, nchan, npol), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _l.check(_l.lib.pfb_inverse_execute_batched(self._h, x.data_ptr(), tmp.data_ptr(), G, npol, n_dat,
                                                        _stream_ptr(self.device)))
        # [T2][g][t1][pol] -> sample index T2*nchan + t1
        return tmp.permute(0, 2, 1, 3).reshape(nblk * keep2 * nchan, G, npol)

    def stage(self, stage, x, out):
        n_dat, nchan, npol = x.shape
        with torch.cuda.device(self.device):
            _l.check(_l.lib.pfb_inverse_execute_stage(self._h, int(stage), x.data_ptr(), out.data_ptr(), npol,
                                                      n_dat, _stream_ptr(self.device)))
        return out


def launch_count():
    return int(_l.lib.pfb_launch_count())


def round_trip(analysis: AnalysisPlan, inverse: InversePlan, x, chan=None, out=None, want_chan=False):
    """Fused analysis -> inversion of a device-resident stream (``pfb_roundtrip_execute``): the analysis kernel writes the
    inversion's channel-IDFT intermediate itself.  Returns (chan or None, y).  Raises PfbError for shapes outside the
    SKA-Low dual-pol fast path (use the two plans separately there)."""
    _check_tensor(x, analysis.dtype, analysis.device, 2)
    n_dat, npol = x.shape
    nb = analysis.nblocks(n_dat)
    if want_chan and chan is None:
        chan = torch.empty((nb, analysis.nchan_out, npol), dtype=analysis.dtype, device=analysis.device)
    if out is None:
        out = torch.empty((inverse.out_len(nb), npol), dtype=analysis.dtype, device=analysis.device)
    with torch.cuda.device(analysis.device):
        _l.check(_l.lib.pfb_roundtrip_execute(analysis._h, inverse._h, x.data_ptr(), chan.data_ptr() if chan is not None else None,
                                              out.data_ptr(), npol, n_dat, _stream_ptr(analysis.device)))
    return chan, out
