"""``synthesize`` with the reference signature (python/data_gen/synthesize.py:27-95)."""
from __future__ import annotations

import logging
import os

from .. import fft_windows
from .. import partialize as _partialize
from ..format_handler import PSRFormatSynthesizer
from ..formats import DADAFile
from . import util
from .config import build_dir, config

__all__ = ["synthesize", "register_backend", "fft_window_lookup"]

module_logger = logging.getLogger(__name__)
_extra_backends = {}

fft_window_lookup = {
    "no_window": lambda a, *args: fft_windows.no_window(a),
    "tukey": fft_windows.tukey_window,
    "top_hat": fft_windows.top_hat_window,
    "hann": lambda a, *args: fft_windows.hann_window(a),
}


def register_backend(name, func):
    """func(input_DADAFile, input_fft_length, input_overlap, fft_window_str, deripple, output_dir, output_file_name) -> DADAFile"""
    _extra_backends[name] = func


@_partialize.partialize
def synthesize(input_data_file_path, input_fft_length: int = None, input_overlap: int = None,
               fft_window_str: str = "no_window", output_file_name: str = None, output_dir: str = "./",
               deripple: bool = True, backend: str = "b200"):
    if input_fft_length is None:
        input_fft_length = config["input_fft_length"]
    if input_overlap is None:
        input_overlap = config["input_overlap"]
    if output_dir is None:
        output_dir = "./"
    matlab_cmd_str = "synthesize"
    output_base = f"{matlab_cmd_str}.{input_fft_length}"
    output_base, log_file_name, output_file_name = util.create_output_file_names(output_file_name, output_base)

    if backend == "matlab":
        cmd_str = (f"{os.path.join(build_dir, matlab_cmd_str)} {input_data_file_path} {input_fft_length} "
                   f"{output_file_name} {output_dir} 1 1 {1 if deripple else 0} {input_overlap} {fft_window_str}")
        util.run_cmd(cmd_str, log_file_path=os.path.join(output_dir, log_file_name))
        return DADAFile(os.path.join(output_dir, output_file_name)).load_data()
    if backend in ("b200", "python"):
        input_data_file = DADAFile(input_data_file_path)
        fft_window = fft_window_lookup[fft_window_str](input_fft_length, input_overlap)
        synthesizer = PSRFormatSynthesizer(input_overlap=input_overlap, fft_window=fft_window,
                                           input_fft_length=input_fft_length, apply_deripple=deripple)
        return synthesizer(input_data_file, output_dir=output_dir, output_file_name=output_file_name)
    if backend in _extra_backends:
        return _extra_backends[backend](DADAFile(input_data_file_path).load_data(), input_fft_length, input_overlap,
                                        fft_window_str, deripple, output_dir, output_file_name)
    raise ValueError(f"unknown backend {backend!r}")
