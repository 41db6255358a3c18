"""Test-vector generators with the reference's definitions and file naming
(python/data_gen/generate_test_vector.py:24-71,74-209)."""
from __future__ import annotations

import os
import typing

import numpy as np

from .. import partialize as _partialize
from ..formats import DADAFile
from . import util
from .config import build_dir, config, config_dir

__all__ = ["complex_sinusoid", "time_domain_impulse", "generate_test_vector"]


def complex_sinusoid(n: int, freqs: typing.List[float], phases: typing.List[float], bin_offset: float = 0.0,
                     dtype: np.dtype = np.complex64):
    """Sum of exp(i (2 pi (f + bin_offset) t / n + phase)); f < 1 means a fraction of n."""
    if not hasattr(freqs, "__iter__"):
        freqs, phases = [freqs], [phases]
    t = np.arange(n)
    sig = np.zeros(n, dtype=dtype)
    for i in range(len(freqs)):
        freq = freqs[i]
        if freq < 1.0:
            freq = int(n * freqs[i])
        sig += np.exp(1j * (2 * np.pi * (freq + bin_offset) / n * t + phases[i])).astype(dtype)
    return sig


def time_domain_impulse(n: int, offsets: typing.List[float], widths: typing.List[int], dtype: np.dtype = np.complex64):
    """sig[o:o+w] = 1; o < 1 means a fraction of n."""
    if not hasattr(offsets, "__iter__"):
        offsets, widths = [offsets], [widths]
    sig = np.zeros(n, dtype=dtype)
    for i in range(len(offsets)):
        offset = offsets[i]
        if offset < 1.0:
            offset = int(offsets[i] * n)
        sig[int(offset): int(offset) + int(widths[i])] = 1.0
    return sig


def _noise(n, dtype=np.float32):
    return (np.random.rand(n) + 1j * np.random.rand(n)).astype(dtype)


@_partialize.partialize
def generate_test_vector(*args, n_bins: int, domain_name: str, header_template: str = None,
                         output_file_name: str = None, output_dir: str = "./", n_pol: int = 1,
                         dtype: np.dtype = np.complex64, backend: str = "python"):
    """``generator = generate_test_vector(backend="python", n_bins=N)`` then
    ``generator([offset], [width], domain_name="time")`` -> DADAFile of shape (n_bins, 1, n_pol)."""
    if output_dir is None:
        output_dir = "./"
    if header_template is None:                       # generate_test_vector.py:121-123 of the reference
        header_template = os.path.join(config_dir, config["header_file_path"])
    if len(args) > 0:
        args_list = []
        for arg in args:
            if hasattr(arg, "__iter__"):
                arg = arg[0]
            args_list.append(f"{arg:.3f}")
        args_str, args_str_comma_sep = "-".join(args_list), ",".join(args_list)
    else:
        args_str = args_str_comma_sep = ""
    matlab_dtype_str = util.matlab_dtype_lookup[dtype]
    output_base = f"{{func_name}}.{n_bins}.{args_str}.{n_pol}.{matlab_dtype_str}.{backend}"

    if backend == "matlab":
        handler = {"time": "time_domain_impulse", "freq": "complex_sinusoid"}[domain_name]
        output_base = output_base.format(func_name=handler)
        output_base, log_file_name, output_file_name = util.create_output_file_names(output_file_name, output_base)
        if header_template is None:
            raise RuntimeError("the matlab backend needs a header template JSON")
        cmd_str = (f"{os.path.join(build_dir, 'generate_test_vector')} {handler} {n_bins} {args_str_comma_sep} "
                   f"{matlab_dtype_str} {n_pol} {header_template} {output_file_name} {output_dir} 1")
        util.run_cmd(cmd_str, log_file_path=os.path.join(output_dir, log_file_name))
        return DADAFile(os.path.join(output_dir, output_file_name)).load_data()

    if backend in ("python", "b200"):
        func_lookup = {"time": time_domain_impulse, "freq": complex_sinusoid, "noise": _noise}
        func = func_lookup[domain_name]
        sig = func(n_bins, *args, dtype=dtype)
        output_data = np.zeros((sig.shape[0], 1, n_pol), dtype=dtype)
        for i_pol in range(n_pol):
            output_data[:, 0, i_pol] = sig
        name = "noise" if domain_name == "noise" else func.__name__
        output_base = output_base.format(func_name=name)
        output_base, log_file_name, output_file_name = util.create_output_file_names(output_file_name, output_base)
        dada_file = DADAFile(os.path.join(output_dir, output_file_name))
        if header_template is not None and os.path.exists(header_template):
            import json
            with open(header_template) as f:
                dada_file.header = json.load(f)
        dada_file.data = output_data
        dada_file.dump_data()
        return dada_file
    raise ValueError(f"unknown backend {backend!r}")
