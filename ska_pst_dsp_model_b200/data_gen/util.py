"""Path / naming helpers with the reference's behaviour (python/data_gen/util.py:77-91,
94-102, 157-186).  The matplotlib-based ``report2plot`` is out of scope."""
from __future__ import annotations

import functools
import os
import shlex
import subprocess

import numpy as np

__all__ = ["updir", "curdir", "run_cmd", "create_output_file_names", "matlab_dtype_lookup", "coro", "rpartial"]

matlab_dtype_lookup = {np.float32: "single", np.float64: "double", np.complex64: "single", np.complex128: "double"}


def run_cmd(cmd_str: str, log_file_path: str = None):
    cmd_split = shlex.split(cmd_str)
    if log_file_path is not None:
        with open(log_file_path, "w") as log_file:
            cmd = subprocess.run(cmd_split, stdout=log_file, stderr=log_file)
    else:
        cmd = subprocess.run(cmd_split)
    if cmd.returncode != 0:
        raise RuntimeError("Exited with non zero status")
    return cmd


def create_output_file_names(output_file_name, default_base):
    if output_file_name is None:
        output_base = default_base
        output_file_name = output_base + ".dump"
    else:
        output_base = os.path.splitext(output_file_name)[0]
    return output_base, output_base + ".log", output_file_name


def updir(base_dir, levels):
    return base_dir if levels == 0 else updir(os.path.dirname(base_dir), levels - 1)


def curdir(file_path):
    return os.path.dirname(os.path.abspath(file_path))


def coro(func):
    @functools.wraps(func)
    def _inner(*args, **kwargs):
        f = func(*args, **kwargs)
        next(f)
        return f
    return _inner


def rpartial(func, *args, **kwargs):
    return lambda *a: func(*(a + args), **kwargs)
