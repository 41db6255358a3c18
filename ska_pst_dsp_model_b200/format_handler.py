"""File-level operators: the ``pfb.format_handler.PSRFormatChannelizer`` /
``PSRFormatSynthesizer`` objects that ``backend="python"`` instantiates in the reference
(python/data_gen/channelize.py:81-91, synthesize.py:84-95), computing on the GPU.

Header handling follows the Matlab executables: channelize.m:79-85 +
add_fir_filter_to_header.m:28-38 (TSAMP, PFB_DC_CHAN, NCHAN_PFB_0, OS_FACTOR, NSTAGE,
COEFF_0 as %0.6E, OVERSAMP_0, NTAP_0) and synthesize.m:76-93 (TSAMP; NCHAN / OS_FACTOR /
COEFF_0 are read from the input header).
"""
from __future__ import annotations

import os

import numpy as np

from .filters import read_fir_filter_coeff, resolve_fir_filter
from .formats import DADAFile
from .rational import Rational


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("the b200 backend needs a CUDA device; there is no CPU fallback")
    return torch


def _tdtype(np_dtype):
    torch = _torch()
    return torch.complex128 if np.dtype(np_dtype) == np.complex128 else torch.complex64


def _stream_it(path) -> bool:
    """Files whose payload exceeds PFB_STREAM_BYTES (default 1 GiB) go through the streaming operators."""
    try:
        limit = int(os.environ.get("PFB_STREAM_BYTES", str(1 << 30)))
        return os.path.getsize(path) > limit
    except OSError:
        return False


def _fmt_num(v: float) -> str:
    s = repr(float(v))
    return s[:-2] if s.endswith(".0") else s


class PSRFormatChannelizer:
    def __init__(self, os_factor, nchan, fir_filter_coeff, use_ifft=True, analysis_function="polyphase_analysis",
                 config=None, config_dir=None):
        self.os_factor = Rational.from_str(os_factor)
        self.nchan = int(nchan)
        self.use_ifft = use_ifft          # accepted for signature compatibility; the maths is fixed
        self.analysis_function = analysis_function
        if isinstance(fir_filter_coeff, (str, os.PathLike)):
            self.fir_filter_coeff = resolve_fir_filter(str(fir_filter_coeff), config, config_dir)
        else:
            self.fir_filter_coeff = np.asarray(fir_filter_coeff, dtype=np.float64).ravel()
        self._plans = {}

    def _plan(self, dtype):
        from . import ops
        if dtype not in self._plans:
            self._plans[dtype] = ops.AnalysisPlan(self.nchan, self.os_factor, self.fir_filter_coeff,
                                                  self.analysis_function, dtype=dtype)
        return self._plans[dtype]

    def channelize_array(self, data):
        """(ndat, 1, npol) complex numpy -> (nblocks, nchan_out, npol)."""
        torch = _torch()
        data = np.asarray(data)
        if data.ndim != 3 or data.shape[1] != 1:
            raise ValueError("expected single-channel input of shape (ndat, 1, npol)")
        if not np.iscomplexobj(data):
            raise ValueError("input data are real-valued")
        cdt = np.complex128 if data.dtype == np.complex128 else np.complex64
        x = torch.from_numpy(np.ascontiguousarray(data[:, 0, :].astype(cdt, copy=False))).cuda()
        out = self._plan(_tdtype(cdt))(x)
        return out.cpu().numpy()

    def __call__(self, input_data_file, output_dir="./", output_file_name=None):
        if input_data_file._data is None and _stream_it(input_data_file.file_path) and \
                self.analysis_function != "polyphase_analysis_padded":
            # large file not yet in memory: stream it piecewise through pinned buffers (file_streaming.py);
            # the returned object holds the header, .load_data() reads the payload
            from .file_streaming import stream_channelize
            return stream_channelize(self, input_data_file.file_path, output_dir, output_file_name)
        if input_data_file._data is None:
            input_data_file.load_data()
        out = self.channelize_array(input_data_file.data)
        if output_file_name is None:
            output_file_name = "channelized." + os.path.basename(input_data_file.file_path)
        of = DADAFile(os.path.join(output_dir, output_file_name), header=input_data_file.header)
        tsamp = float(input_data_file.header.get("TSAMP", "1"))
        of["TSAMP"] = _fmt_num(tsamp * self.nchan * self.os_factor.de / self.os_factor.nu)   # channelize.m:81
        of["PFB_DC_CHAN"] = "1"
        of["NCHAN_PFB_0"] = str(self.nchan)
        of["OS_FACTOR"] = str(self.os_factor)
        of["NSTAGE"] = "1"
        of["COEFF_0"] = ",".join("%0.6E" % c for c in self.fir_filter_coeff)
        of["OVERSAMP_0"] = str(self.os_factor)
        of["NTAP_0"] = str(self.fir_filter_coeff.size)
        of.data = out
        of.dump_data()
        return of


class PSRFormatSynthesizer:
    def __init__(self, input_overlap, fft_window, input_fft_length, apply_deripple=True, spans_nyquist=True, combine=1):
        self.input_overlap = int(input_overlap)
        self.input_fft_length = int(input_fft_length)
        self.fft_window = np.asarray(fft_window, dtype=np.float64).ravel()
        self.apply_deripple = bool(apply_deripple)
        self.spans_nyquist, self.combine = spans_nyquist, combine
        self._plans = {}

    def _plan(self, nchan, os_factor, taps, dtype):
        from . import ops
        key = (nchan, str(os_factor), None if taps is None else taps.tobytes(), dtype)
        if key not in self._plans:
            self._plans = {key: ops.InversePlan(nchan, os_factor, self.input_fft_length, self.input_overlap,
                                                self.fft_window, deripple_taps=taps, spans_nyquist=self.spans_nyquist,
                                                combine=self.combine, dtype=dtype)}
        return self._plans[key]

    def synthesize_array(self, data, os_factor, fir_filter_coeff=None, sample_offset=1):
        """(ndat, nchan, npol) complex numpy -> (nout, 1, npol)."""
        torch = _torch()
        data = np.asarray(data)
        if not np.iscomplexobj(data):
            raise ValueError("polyphase_synthesis input data are real-valued!")
        cdt = np.complex128 if data.dtype == np.complex128 else np.complex64
        taps = None
        if self.apply_deripple:
            if fir_filter_coeff is None:
                raise ValueError("deripple requested but no filter coefficients available (COEFF_0 missing)")
            taps = np.asarray(fir_filter_coeff, dtype=np.float64).ravel()
        plan = self._plan(data.shape[1], Rational.from_str(os_factor), taps, _tdtype(cdt))
        x = torch.from_numpy(np.ascontiguousarray(data.astype(cdt, copy=False))).cuda()
        out = plan(x, sample_offset=sample_offset)
        return out.cpu().numpy()[:, None, :]

    def __call__(self, input_data_file, output_dir="./", output_file_name=None):
        if input_data_file._data is None and _stream_it(input_data_file.file_path):
            from .file_streaming import stream_synthesize
            return stream_synthesize(self, input_data_file.file_path, output_dir, output_file_name)
        if input_data_file._data is None:
            input_data_file.load_data()
        hdr = input_data_file.header
        os_factor = Rational.from_str(hdr["OS_FACTOR"])
        taps = None
        if self.apply_deripple and "COEFF_0" in hdr:
            taps = np.array([float(v) for v in hdr["COEFF_0"].split(",")])
        out = self.synthesize_array(input_data_file.data, os_factor, taps)
        if output_file_name is None:
            output_file_name = "synthesized." + os.path.basename(input_data_file.file_path)
        of = DADAFile(os.path.join(output_dir, output_file_name), header=hdr)
        nchan = input_data_file.data.shape[1]
        tsamp = float(hdr.get("TSAMP", "1"))
        of["TSAMP"] = _fmt_num(tsamp * os_factor.nu / os_factor.de / nchan)
        of.data = out
        of.dump_data()
        return of
