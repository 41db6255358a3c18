"""Temporal tapers of the inversion: ``pfb.fft_windows.{no_window, tukey_window,
top_hat_window, hann_window}`` as called by python/data_gen/synthesize.py:19-24,
defined as in matlab/PFBWindow.m (18-23, 26-42, 59-66, 68-100)."""
from __future__ import annotations

import numpy as np


def _hann_sym(n):
    """Matlab ``hann(n)``: symmetric, zero end points."""
    if n == 1:
        return np.ones(1)
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(n) / (n - 1)))


def no_window(n, *_):
    return np.ones(int(n))


def tukey_window(n, overlap):
    w = np.ones(int(n))
    ov = int(overlap)
    if ov > 0:
        h = _hann_sym(2 * ov)
        w[:ov] = h[:ov]
        w[n - ov:] = h[ov:]
    return w


def top_hat_window(n, overlap):
    w = np.ones(int(n))
    ov = int(overlap)
    if ov > 0:
        w[:ov] = 0.0
        w[n - ov:] = 0.0
    return w


def hann_window(n, *_):
    return _hann_sym(int(n))


fft_window_lookup = {
    "no_window": no_window,
    "tukey": tukey_window,
    "top_hat": top_hat_window,
    "hann": hann_window,
}


def make_window(name, n, overlap):
    if name not in fft_window_lookup:
        raise KeyError(f"unknown fft window {name!r}; choose from {sorted(fft_window_lookup)}")
    return fft_window_lookup[name](n, overlap)
