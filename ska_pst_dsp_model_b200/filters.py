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


def interpft(x, ny):
    """Matlab's `interpft(x, ny)` for a real vector: resampling of the trigonometric interpolant through the FFT (zero
    padding in the middle of the spectrum, the Nyquist bin of an even-length input split in two, scale ny/m); when
    ny < m Matlab interpolates to the next multiple incr*ny > m and keeps every incr-th sample."""
    x = np.asarray(x, dtype=np.float64)
    m = x.size
    incr = 1
    if ny <= m:
        incr = m // ny + 1
        ny = incr * ny
    a = np.fft.fft(x)
    nyq = (m + 2) // 2                                   # ceil((m + 1) / 2)
    b = np.concatenate([a[:nyq], np.zeros(ny - m, dtype=complex), a[nyq:]])
    if m % 2 == 0:
        b[nyq - 1] /= 2
        b[nyq - 1 + ny - m] = b[nyq - 1]
    y = np.real(np.fft.ifft(b)) * ny / m
    return y[::incr]


def maxflat_sym(order, wn):
    """Symmetric (linear-phase) maximally flat FIR low-pass of even `order` whose magnitude passes 1/sqrt(2) nearest to
    `wn` (units of Nyquist): Herrmann's closed form A(w) = cos^(2L)(w/2) * sum_{k<K} C(L-1+k, k) sin^(2k)(w/2) with L
    zeros at z = -1 and K = order/2 + 1 - L degrees of flatness at w = 0 - the filter Matlab's `maxflat(n, 'sym', wn)`
    designs (generate_MaxFlt.m:35).  Unpinned against Matlab (not available): L is the integer whose half-power
    frequency is closest to wn."""
    from math import comb
    if order % 2:
        raise ValueError("maxflat_sym: the order must be even")
    M = order // 2

    def response(L, w):
        c2, s2 = np.cos(w / 2) ** 2, np.sin(w / 2) ** 2
        return c2 ** L * sum(comb(L - 1 + k, k) * s2 ** k for k in range(M + 1 - L))

    grid = np.linspace(0, np.pi, 8193)
    best = None
    for L in range(1, M + 1):
        r = response(L, grid)
        wc = grid[np.argmin(np.abs(r - np.sqrt(0.5)))] / np.pi
        if best is None or abs(wc - wn) < best[0]:
            best = (abs(wc - wn), L)
    L = best[1]
    n = order + 1
    A = response(L, 2 * np.pi * np.arange(n) / n)        # a cosine polynomial of degree M: n samples determine it
    h = np.real(np.fft.ifft(A))
    return np.roll(h, M)                                 # centred: h[M] is the peak


def generate_maxflat(nbuff, ntap, iterations=10):
    """The maximally flat prototype of matlab/generate_MaxFlt.m:32-71 (J. Bunton): a 2*ntap-tap two-channel filter from
    `maxflat`, ten fixed-point passes that push |H(f)|^2 + |H(f + pi)|^2 towards 1 (constant total power over the
    channels for a monochromatic input), then trigonometric interpolation to nbuff*ntap taps.  Returns (W, residual)
    with residual = max |1 - (|H|^2 + |H shifted|^2) / value at DC| of the 2*ntap-tap filter."""
    n2 = 2 * ntap
    imp = maxflat_sym(n2, 0.5 * n2 / (n2 + 1))
    imp = interpft(imp, n2) * (n2 + 1) / n2
    alt = np.cos(np.arange(n2) * np.pi)
    resid = None
    for _ in range(iterations + 1):
        f, fh = np.fft.fft(imp), np.fft.fft(imp * alt)
        e = np.real(f * np.conj(f) + fh * np.conj(fh))
        e = 1.0 - e / e[0]
        resid = float(np.abs(e).max())
        if _ == iterations:
            break
        imp = imp + np.real(np.fft.fftshift(np.fft.ifft(e))) / 2.0
    return interpft(imp, nbuff * ntap), resid


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
