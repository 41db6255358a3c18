"""``channelize`` with the reference signature (python/data_gen/channelize.py:19-92).

backend="b200" (and "python", so that verify.test_backends / verify.purity drive it
unchanged) runs the CUDA analysis; backend="matlab" still shells out to the mcc
executable ``build/channelize`` with the reference's argv (channelize.py:66-69).
Extra backends can be registered for cross-implementation tests."""
from __future__ import annotations

import logging
import os

from .. import partialize as _partialize
from ..format_handler import PSRFormatChannelizer
from ..formats import DADAFile
from . import util
from .config import build_dir, config, config_dir

__all__ = ["channelize", "register_backend"]

module_logger = logging.getLogger(__name__)
_extra_backends = {}


def register_backend(name, func):
    """func(input_DADAFile, channels, os_factor_str, fir_filter_path, output_dir, output_file_name, use_padded) -> DADAFile"""
    _extra_backends[name] = func


@_partialize.partialize
def channelize(input_data_file_path: str, channels: int = None, os_factor_str: str = None,
               fir_filter_path: str = None, output_file_name: str = None, output_dir: str = "./",
               backend: str = "b200", use_padded: bool = False):
    if channels is None:
        channels = config["channels"]
    if os_factor_str is None:
        os_factor_str = config["os_factor"]
    os_factor_str = str(os_factor_str)
    if fir_filter_path is None:
        fir_filter_path = os.path.join(config_dir, config["fir_filter_coeff_file_path"])
    if output_dir is None:
        output_dir = "./"
    module_logger.debug(f"channelize: {input_data_file_path} channels={channels} os={os_factor_str} "
                        f"fir={fir_filter_path} out={output_dir}/{output_file_name} backend={backend}")
    matlab_cmd_str = "channelize"
    output_base = f"{matlab_cmd_str}.{channels}.{'-'.join(os_factor_str.split('/'))}"
    output_base, log_file_name, output_file_name = util.create_output_file_names(output_file_name, output_base)

    if backend == "matlab":
        cmd_str = (f"{os.path.join(build_dir, matlab_cmd_str)} {input_data_file_path} {channels} {os_factor_str} "
                   f"{fir_filter_path} {output_file_name} {output_dir} 1 {1 if use_padded else 0}")
        util.run_cmd(cmd_str, log_file_path=os.path.join(output_dir, log_file_name))
        return DADAFile(os.path.join(output_dir, output_file_name)).load_data()
    if backend in ("b200", "python"):
        analysis = config.get("analysis_function", "polyphase_analysis")
        if use_padded:
            analysis = "polyphase_analysis_padded"
        elif analysis == "polyphase_analysis_padded" and channels != config.get("channels"):
            analysis = "polyphase_analysis"
        input_data_file = DADAFile(input_data_file_path)
        channelizer = PSRFormatChannelizer(use_ifft=True, os_factor=os_factor_str, nchan=channels,
                                           fir_filter_coeff=fir_filter_path, analysis_function=analysis,
                                           config=config, config_dir=config_dir)
        return channelizer(input_data_file, output_dir=output_dir, output_file_name=output_file_name)
    if backend in _extra_backends:
        return _extra_backends[backend](DADAFile(input_data_file_path).load_data(), channels, os_factor_str,
                                        fir_filter_path, output_dir, output_file_name, use_padded)
    raise ValueError(f"unknown backend {backend!r}")
