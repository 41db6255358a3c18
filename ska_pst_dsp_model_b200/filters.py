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


# ---------------------------------------------------------------------------
# Prototype-filter designers.  The .mat files named by the reference configs are
# git-ignored upstream, so they are re-created here with scipy.signal.firls
# following matlab/design_PFB_FIR_filter.m:24-48 and
# matlab/design_PFB_FIR_filter_two_stage.m:42-78.
# ---------------------------------------------------------------------------
def design_pfb_fir_filter(n_chan, os_nu, os_de, n_taps_per_chan):
    """firls, order n_chan*taps_per_chan, bands [0, 1/N, (2 OS - 1)/N, 1], weights [1, 15]."""
    from scipy.signal import firls
    os_f = os_nu / os_de
    if os_f == 1:
        os_f += 0.1
    fp, fs = 1.0 / n_chan, (2 * os_f - 1) / n_chan
    return firls(n_chan * n_taps_per_chan + 1, [0, fp, fs, 1.0], [1, 1, 0, 0], weight=[1, 15])


def design_pfb_fir_filter_two_stage(n_chan, os_nu, os_de, os_taps_per_chan=28, zero_stuff_factor=32):
    """Stage-1 firls design at ``zero_stuff_factor`` x the band edges (0.998 Fs), then
    FFT-domain zero stuffing (interpolation) to the full length (100 353 taps for Mid)."""
    from scipy.signal import firls
    os_f = os_nu / os_de
    n_taps = int(round(os_taps_per_chan * n_chan / os_f))
    n1 = n_taps // zero_stuff_factor
    fp, fs = 1.0 / n_chan, (2 * os_f - 1) / n_chan
    h0 = firls(n1 + 1, [0, fp * zero_stuff_factor, 0.998 * fs * zero_stuff_factor, 1.0], [1, 1, 0, 0], weight=[1, 15])
    h1 = np.fft.fft(np.fft.ifftshift(h0))
    hz = np.concatenate([h1[: n1 // 2 + 1], np.zeros(n1 * (zero_stuff_factor - 1)), h1[n1 // 2 + 1:]])
    return np.real(np.fft.fftshift(np.fft.ifft(hz)))


_DESIGN_CACHE = {}


def resolve_fir_filter(path, config=None, config_dir=None):
    """Coefficients for a config entry: an existing file wins; the shipped PST taps are
    recognised by name; git-ignored upstream designs are re-created (and cached)."""
    candidates = [path]
    if config_dir is not None and not os.path.isabs(path):
        candidates.append(os.path.join(config_dir, path))
    for c in candidates:
        if os.path.exists(c):
            return read_fir_filter_coeff(c)
    base = os.path.basename(path)
    if base == "PST_filtertaps.txt":
        return load_pst_taps()
    design = (config or {}).get("fir_filter_design")
    if design is None:
        raise FileNotFoundError(f"FIR coefficient file {path} not found and no 'fir_filter_design' to re-create it")
    nu, de = (int(v) for v in str(config["os_factor"]).split("/"))
    key = (design, config["channels"], nu, de, config["fir_filter_taps"])
    if key not in _DESIGN_CACHE:
        import logging
        logging.getLogger(__name__).warning(
            "FIR coefficient file %s not found: re-creating the %s design with scipy.signal.firls "
            "(coefficients will differ in the last digits from the Matlab firls design)", path, design)
        cache = os.path.join(_DATA, "designed", f"{design}.{nu}-{de}.{config['channels']}.{config['fir_filter_taps']}.npy")
        if os.path.exists(cache):
            _DESIGN_CACHE[key] = np.load(cache)
        else:
            if design == "single_stage":
                h = design_pfb_fir_filter(config["channels"], nu, de, (config["fir_filter_taps"] - 1) // config["channels"]
                                          if config["fir_filter_taps"] % config["channels"] else config["fir_filter_taps"] // config["channels"])
            elif design == "two_stage":
                h = design_pfb_fir_filter_two_stage(config["channels"], nu, de)
            else:
                raise ValueError(f"unknown fir_filter_design {design!r}")
            try:
                os.makedirs(os.path.dirname(cache), exist_ok=True)
                np.save(cache, h)
            except OSError:
                pass
            _DESIGN_CACHE[key] = h
    return _DESIGN_CACHE[key]
