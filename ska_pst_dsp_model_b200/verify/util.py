"""Spurious-power metrics of the purity report (python/verify/util.py:15-43)."""
import numpy as np

__all__ = ["spurious", "total_spurious", "mean_spurious", "max_spurious", "dB"]


def spurious(a):
    b = a.copy()
    b[np.argmax(b)] = 0.0
    return b


def dB(a):
    """``a`` is already a power."""
    return 10.0 * np.log10(np.abs(a) + 1e-13)


def total_spurious(a):
    return dB(np.sum(spurious(np.abs(a) ** 2)))


def mean_spurious(a):
    return dB(np.mean(spurious(np.abs(a) ** 2)))


def max_spurious(a):
    return dB(np.amax(spurious(np.abs(a) ** 2)))
