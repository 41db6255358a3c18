"""Prototype-filter coefficient I/O (matlab/read_fir_filter_coeff.m:11-18) and the
shipped SKA-Low PST taps."""
from __future__ import annotations

import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
PST_TAPS_PATH = os.path.join(_DATA, "pst_lowcbf_taps_int.txt")


def load_pst_taps(normalise: bool = False) -> np.ndarray:
    """The 3072 integer taps of the reference's config/PST_filtertaps.txt
    (sum 16 777 241 ~ 2^24, symmetric about index 1536)."""
    t = np.loadtxt(PST_TAPS_PATH, dtype=np.float64)
    return t / t.sum() if normalise else t


def read_fir_filter_coeff(file_path: str) -> np.ndarray:
    """``.mat`` with an ``hQ`` or ``h`` field, or a plain text table of numbers
    (read_fir_filter_coeff.m:11-18: hQ wins over h; anything else is taken as
    the array itself)."""
    ext = os.path.splitext(file_path)[1].lower()
    if ext == ".mat":
        from scipy.io import loadmat
        m = loadmat(file_path)
        for key in ("hQ", "h"):
            if key in m:
                return np.asarray(m[key], dtype=np.float64).ravel()
        arrays = [v for k, v in m.items() if not k.startswith("__")]
        if len(arrays) != 1:
            raise ValueError(f"{file_path}: expected an 'h' or 'hQ' field")
        return np.asarray(arrays[0], dtype=np.float64).ravel()
    if ext == ".npy":
        return np.load(file_path).astype(np.float64).ravel()
    return np.loadtxt(file_path, dtype=np.float64).ravel()
