"""CPU oracle for the PFB analysis + Golden inversion hot path.

TEST INFRASTRUCTURE ONLY.  This module is the parity checker for the CUDA
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``ska_pst_dsp_model_b200/`` imports it, and the product path raises when the
CUDA library is missing rather than falling back to this code.

It restates, in numpy, the Matlab "Golden" signal model of
ska-telescope/ska-pst-dsp-model (the reference).  Every function names the
reference file:line it follows.  Each hot function exists in two forms:

* ``*_literal``  – a line-by-line transliteration of the Matlab loop
  (1-based indices converted, same order of operations);
* the un-suffixed fast form – the vectorised closed form, used
  for larger parity cases and as the timed CPU baseline.

``tests/test_oracle.py`` checks literal == closed form to <=1e-12.

PINNING STATUS: the reference ships no golden vectors (``data/`` is empty,
``*.dump`` is git-ignored) and neither Matlab/Octave nor the third-party
``pfb`` package can run in the build container, so array-level parity with the
reference executables is **unpinned**.  The oracle is pinned on every
known-answer / behavioural criterion the reference's own tests hold for this
path: size formulas (``calc_output_nbins.m``,
``test_dspsr_util.py:77-81``), the TestImpulse / TestPureTone /
TestFrequencyComb pass criteria, the chop/alignment rule, header behaviour and
the shipped-taps invariants.  See tests/test_oracle.py and
tests/test_oracle_pins.py.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

__all__ = [
    "OSFactor", "pad_filter", "normalize", "multiply", "calc_output_nbins",
    "polyphase_analysis", "polyphase_analysis_literal",
    "polyphase_analysis_padded", "polyphase_analysis_padded_literal",
    "pst_filterbank_literal", "polyphase_analysis_lowcbf",
    "hann", "window_no_window", "window_tukey", "window_top_hat",
    "spectral_hann", "make_temporal_taper",
    "freqz_half", "deripple_response", "combine_channel_map",
    "polyphase_synthesis", "polyphase_synthesis_literal",
    "FilterBank", "InverseFilterBank", "TwoStageFilterBank",
    "TwoStageInverseFilterBank",
    "complex_sinusoid", "time_domain_impulse", "pure_tone", "frequency_comb",
    "frequency_comb_params", "impulse",
    "design_pfb_fir_filter", "design_pfb_fir_filter_two_stage",
]


# --------------------------------------------------------------------------
# rational oversampling factor + small helpers
# --------------------------------------------------------------------------
class OSFactor:
    """struct('nu',..,'de',..) of the Matlab code; ``pfb.rational.Rational``
    of the Python harness (``verify/test_backends.py:51-54``)."""

    def __init__(self, nu: int, de: int):
        self.nu = int(nu)
        self.de = int(de)

    @classmethod
    def from_str(cls, s) -> "OSFactor":
        if isinstance(s, OSFactor):
            return s
        nu, de = str(s).split("/")
        return cls(int(nu), int(de))

    def normalize(self, n):
        """matlab/normalize.m:17  (de*n)/nu"""
        v = Fraction(self.de * n, self.nu)
        return int(v) if v.denominator == 1 else float(v)

    def multiply(self, n):
        """matlab/multiply.m:17  (nu*n)/de"""
        v = Fraction(self.nu * n, self.de)
        return int(v) if v.denominator == 1 else float(v)

    def __float__(self):
        return self.nu / self.de

    def __str__(self):
        return f"{self.nu}/{self.de}"

    __repr__ = __str__


def _os(os_factor) -> OSFactor:
    if isinstance(os_factor, OSFactor):
        return os_factor
    if isinstance(os_factor, str):
        return OSFactor.from_str(os_factor)
    if isinstance(os_factor, dict):
        return OSFactor(os_factor["nu"], os_factor["de"])
    if isinstance(os_factor, (tuple, list)):
        return OSFactor(*os_factor)
    return OSFactor(os_factor.nu, os_factor.de)


def normalize(os_factor, n):
    return _os(os_factor).normalize(n)


def multiply(os_factor, n):
    return _os(os_factor).multiply(n)


def pad_filter(filter_coeff, n_chan):
    """matlab/pad_filter.m:10-12 – zero-pad taps to a multiple of n_chan."""
    filter_coeff = np.asarray(filter_coeff).ravel()
    phases = int(math.ceil(filter_coeff.size / n_chan))
    padded = np.zeros(phases * n_chan, dtype=filter_coeff.dtype)
    padded[: filter_coeff.size] = filter_coeff
    return padded


def calc_output_nbins(nbins, channels, os_factor, filter_taps,
                      input_fft_length, input_overlap):
    """matlab/calc_output_nbins.m:17-27."""
    os_factor = _os(os_factor)
    step = (channels * os_factor.de) // os_factor.nu
    nblocks_pfb = (nbins - filter_taps) // step
    output_pfb = (step * nblocks_pfb) // channels
    input_keep = input_fft_length - 2 * input_overlap
    nblocks_ipfb = (output_pfb - 2 * input_overlap) // input_keep
    output_fft_length = os_factor.normalize(input_fft_length) * channels
    output_overlap = os_factor.normalize(input_overlap) * channels
    output_keep = output_fft_length - 2 * output_overlap
    return output_keep * nblocks_ipfb


def _real_dtype(dtype):
    return np.float32 if np.dtype(dtype) in (np.dtype(np.complex64), np.dtype(np.float32)) else np.float64


def _cplx_dtype(dtype):
    return np.complex64 if _real_dtype(dtype) is np.float32 else np.complex128


# --------------------------------------------------------------------------
# analysis: polyphase_analysis.m  (Bunton, SKA-Low)
# --------------------------------------------------------------------------
def polyphase_analysis_literal(x, filt, block, os_factor):
    """matlab/polyphase_analysis.m:46-122, loop for loop.

    x: (n_pol, 1, n_dat) complex; returns (n_pol, block, nblocks).
    Arithmetic in the class of ``x`` (``cast(filt, class(in))`` :53).
    """
    os_factor = _os(os_factor)
    x = np.asarray(x)
    n_pol, _, n_dat = x.shape
    cdt = _cplx_dtype(x.dtype)
    filt = np.asarray(filt).ravel().astype(_real_dtype(x.dtype))
    step = (block * os_factor.de) // os_factor.nu            # :56
    f = pad_filter(filt, block)                               # :58
    phases = f.size // block                                  # :60
    fl = f.size
    nblocks = (n_dat - fl) // step                            # :62
    out = np.zeros((n_pol, block, max(nblocks, 0)), dtype=cdt)
    for i_pol in range(n_pol):                                # :83
        in_pol = x[i_pol, 0, :].astype(cdt)
        for k in range(nblocks):                              # :88
            in_block = in_pol[step * k: fl + step * k]       # :97
            index = step * k - (step * k // block) * block    # :102
            # :105 circshift(f.*in_block, index)'  -- ' is the CONJUGATE transpose
            temp = np.conj(np.roll(f * in_block, index))
            temp2 = np.zeros(block, dtype=cdt)                # :112
            for m in range(phases):                           # :113-115
                temp2 = temp2 + temp[block * m: block * (m + 1)]
            # :120 conj(ifft(temp2)*block^2)
            out[i_pol, :, k] = np.conj(np.fft.ifft(temp2).astype(cdt) * (block ** 2))
    return out


def _analysis_core(x1d, f, block, step, nblocks, start0, rot_extra, scale, cdt, chunk=None):
    """out[c,k] = scale * sum_n f[n] x[t0_k+n] W^{c (t0_k+n+rot_extra)},
    t0_k = start0 + k*step, W = exp(-2i pi/block), x = 0 outside the array.
    Closed form of SURVEY §7 (all three analysis variants reduce to it)."""
    fl = f.size
    phases = fl // block
    out = np.empty((block, nblocks), dtype=cdt)
    if nblocks <= 0:
        return out
    lo = start0
    hi = start0 + (nblocks - 1) * step + fl
    pad_lo, pad_hi = max(0, -lo), max(0, hi - x1d.size)
    if pad_lo or pad_hi:
        x1d = np.concatenate([np.zeros(pad_lo, x1d.dtype), x1d, np.zeros(pad_hi, x1d.dtype)])
    base = start0 + pad_lo
    if chunk is None:
        chunk = max(1, (1 << 22) // fl)
    st = x1d.strides[0]
    for k0 in range(0, nblocks, chunk):
        kb = min(chunk, nblocks - k0)
        fr = np.lib.stride_tricks.as_strided(x1d[base + k0 * step:], shape=(kb, fl),
                                             strides=(step * st, st), writeable=False)
        y = (fr * f).reshape(kb, phases, block).sum(axis=1)          # polyphase arms
        sp = np.fft.fft(y, axis=1)
        t0 = (start0 + (k0 + np.arange(kb, dtype=np.int64)) * step + rot_extra) % block
        ut, inv = np.unique(t0, return_inverse=True)     # only nu distinct rotations occur
        ph = np.exp(-2j * np.pi * np.outer(ut, np.arange(block)) / block)[inv]
        out[:, k0:k0 + kb] = (scale * sp * ph).T.astype(cdt)
    return out


def polyphase_analysis(x, filt, block, os_factor):
    """Closed form of matlab/polyphase_analysis.m: upper-sideband
    forward DFT with absolute-time phase reference, x N."""
    os_factor = _os(os_factor)
    x = np.asarray(x)
    n_pol, _, n_dat = x.shape
    cdt = _cplx_dtype(x.dtype)
    f = pad_filter(np.asarray(filt).ravel().astype(_real_dtype(x.dtype)), block)
    step = (block * os_factor.de) // os_factor.nu
    nblocks = max((n_dat - f.size) // step, 0)
    out = np.zeros((n_pol, block, nblocks), dtype=cdt)
    for p in range(n_pol):
        out[p] = _analysis_core(np.ascontiguousarray(x[p, 0, :]).astype(cdt), f, block, step,
                                nblocks, 0, 0, block, cdt)
    return out


# --------------------------------------------------------------------------
# analysis: polyphase_analysis_padded.m  (Gunaratne, SKA-Mid)
# --------------------------------------------------------------------------
def polyphase_analysis_padded_literal(x, filt, block, os_factor):
    """matlab/polyphase_analysis_padded.m:49-156, loop for loop.

    The Matlab allocates ``out`` as real double (:104) so the result is always
    double complex regardless of the input class."""
    os_factor = _os(os_factor)
    x = np.asarray(x)
    n_pol, _, n_dat = x.shape
    filt = np.asarray(filt).ravel().astype(_real_dtype(x.dtype))   # :56
    step = (block * os_factor.de) // os_factor.nu                   # :61
    overlap = block - step                                          # :62
    nblocks = n_dat // step                                         # :75
    sample_delay_shift = int(math.ceil((filt.size - 1) / 2 / step))  # :89
    filt_padded = pad_filter(filt, block)                           # :97
    phases = filt_padded.size // block
    # reshape(filt_padded, block, phases): column-major -> [j, m] = f[j + m*block]
    filt_2d = filt_padded.reshape(phases, block).T
    out = np.zeros((n_pol, block, nblocks), dtype=np.complex128)
    for i_pol in range(n_pol):
        in_pol = x[i_pol, 0, :]
        in_mask = np.zeros(block * phases, dtype=np.result_type(in_pol.dtype, np.complex64))
        in_mask_2d = in_mask.reshape(phases, block).T
        bri = 0
        for idx in range(1, nblocks + 1):                           # :113
            ypfb = (filt_2d * in_mask_2d).sum(axis=1)               # :118
            in_mask[step:] = in_mask[:-step].copy()                 # :121
            in_mask[:step] = in_pol[(idx - 1) * step: idx * step][::-1]  # :124
            in_mask_2d = in_mask.reshape(phases, block).T           # :126
            if bri == 0:
                y1s = ypfb
            else:
                index = ((os_factor.nu - bri) * overlap) % block    # :137
                y1s = np.roll(ypfb, -index)                         # :143
            out[i_pol, :, idx - 1] = (block ** 2) * np.fft.ifft(y1s)  # :147
            bri = (bri + 1) % os_factor.nu                          # :151
    return np.roll(out, -sample_delay_shift, axis=2)                # :156


def polyphase_analysis_padded(x, filt, block, os_factor, out_dtype=None):
    """Closed form of polyphase_analysis_padded.m:
    out[c,i] = N e^{-2 pi i c/N} sum_{t<T} f[T-1-t] x[t] W^{ct},  T = i*S,
    zero initial state, then a circular shift by -ceil((L-1)/(2S)) blocks."""
    os_factor = _os(os_factor)
    x = np.asarray(x)
    n_pol, _, n_dat = x.shape
    cdt = np.complex128 if out_dtype is None else _cplx_dtype(out_dtype)
    filt = np.asarray(filt).ravel()
    f = pad_filter(filt.astype(np.float64 if cdt is np.complex128 else np.float32), block)
    fl = f.size
    step = (block * os_factor.de) // os_factor.nu
    nblocks = n_dat // step
    sds = int(math.ceil((filt.size - 1) / 2 / step))
    fr = f[::-1].copy()                     # f[T-1-t] = fr[t-(T-fl)]
    out = np.zeros((n_pol, block, nblocks), dtype=cdt)
    for p in range(n_pol):
        o = _analysis_core(np.ascontiguousarray(x[p, 0, :]).astype(cdt), fr, block, step,
                           nblocks, -fl, 1, block, cdt)
        out[p] = np.roll(o, -sds, axis=1) if nblocks else o
    return out


# --------------------------------------------------------------------------
# analysis: PSTFilterbank.m + polyphase_analysis_lowcbf.m (Low-CBF firmware)
# --------------------------------------------------------------------------
def pst_filterbank_literal(din, fir_taps, do_padding=True):
    """matlab/PSTFilterbank.m:1-45, loop for loop. din: 1-D complex."""
    din = np.asarray(din).ravel()
    fir_taps = np.asarray(fir_taps, dtype=np.float64).ravel()
    nfilt = 3072
    padding = 1536 if do_padding else 0
    total = din.size + padding
    n_out = (total - nfilt) // 192
    dinp = np.zeros(total, dtype=np.complex128)
    dinp[padding:] = din
    dout = np.zeros((216, max(n_out, 0)), dtype=np.complex128)
    fft_in = np.zeros(256, dtype=np.complex128)
    freq = np.arange(-128, 128)
    for o in range(n_out):
        for n1 in range(256):
            fft_in[n1] = np.sum(fir_taps[n1::256] * dinp[o * 192 + n1: o * 192 + n1 + 256 * 12: 256]) / 2 ** 9
        dout1 = np.fft.fftshift(np.fft.fft(fft_in)) / 128
        rotation = np.mod(o * freq, 4)
        dout2 = dout1 * np.exp(1j * 2 * np.pi * rotation / 4)
        dout[:, o] = dout2[20:236]
    return dout


def polyphase_analysis_lowcbf(x, filt, block=256, os_factor=None, do_padding=True, literal=False):
    """matlab/polyphase_analysis_lowcbf.m:12-47.  ``do_padding`` models the
    Matlab ``persistent`` flag (true on the first call of a session only).
    Closed form: Bunton analysis (N=256, S=192) of the zero-pre-padded input,
    gain 4096 instead of N, FFT bins (ci+148) mod 256 for ci = 0..215."""
    x = np.asarray(x)
    n_pol = x.shape[0]
    scale = 2 ** 9 * 2048 * 256
    outs = []
    for p in range(n_pol):
        if literal:
            outs.append(pst_filterbank_literal(x[p, 0, :], filt, do_padding) * scale)
        else:
            f = np.asarray(filt, dtype=np.float64).ravel()
            assert f.size == 3072
            pad = 1536 if do_padding else 0
            n_out = max((x.shape[2] + pad - 3072) // 192, 0)
            full = _analysis_core(np.ascontiguousarray(x[p, 0, :]).astype(np.complex128), f, 256, 192,
                                  n_out, -pad, pad, 4096.0, np.complex128)
            outs.append(full[(np.arange(216) + 148) % 256, :])
    return np.stack(outs, axis=0)


# --------------------------------------------------------------------------
# tapers: PFBWindow.m
# --------------------------------------------------------------------------
def hann(L):
    """Matlab ``hann(L)`` (symmetric, zero end points)."""
    if L == 1:
        return np.ones(1)
    n = np.arange(L)
    return 0.5 * (1 - np.cos(2 * np.pi * n / (L - 1)))


def window_no_window(n, *_):
    """PFBWindow.m:18-23"""
    return np.ones(n)


def window_tukey(n, discard):
    """PFBWindow.m:26-42: ones with hann(2*discard) halves at both ends."""
    w = np.ones(n)
    if discard > 0:
        h = hann(2 * discard)
        w[:discard] = h[:discard]
        w[n - discard:] = h[discard:]
    return w


def window_top_hat(n, discard):
    """PFBWindow.m:59-66"""
    w = np.ones(n)
    if discard > 0:
        w[:discard] = 0
        w[n - discard:] = 0
    return w


def spectral_hann(length):
    """PFBWindow.m:68-100 used as *spectral* taper on the stitched column
    vector: the window is rebuilt as circshift(hann(len), len/2)."""
    return np.roll(hann(length), length // 2)


def make_temporal_taper(name, n, discard):
    return {"no_window": window_no_window, "tukey": window_tukey,
            "top_hat": window_top_hat, None: window_no_window}[name](n, discard)


# --------------------------------------------------------------------------
# deripple response: polyphase_synthesis.m:138-150
# --------------------------------------------------------------------------
def freqz_half(b, n_points, n_eval):
    """First ``n_eval`` points of Matlab ``freqz(b, 1, n_points)``:
    H(e^{jw}), w = pi*k/n_points."""
    b = np.asarray(b, dtype=np.float64).ravel()
    k = np.arange(n_eval)
    out = np.empty(n_eval, dtype=np.complex128)
    m = np.arange(b.size)
    for i in range(0, n_eval, 64):
        kk = k[i:i + 64]
        out[i:i + 64] = np.exp(-1j * np.pi * np.outer(kk, m) / n_points) @ b
    return out


def deripple_response(filter_coeff, n_chan, fn_width):
    """filter_response (length FN_width/2+1) of polyphase_synthesis.m:142-149."""
    pb = fn_width // 2
    h0 = freqz_half(filter_coeff, n_chan * pb, pb + 1)
    return 1.0 / np.abs(h0)


def deripple_fn_weights(filter_coeff, n_chan, fn_width):
    """Per-bin weight in FN order (polyphase_synthesis.m:244-250):
    FN[i] *= R[pb-i], FN[pb+i] *= R[i], i in [0,pb)."""
    pb = fn_width // 2
    r = deripple_response(filter_coeff, n_chan, fn_width)
    w = np.empty(fn_width)
    i = np.arange(pb)
    w[i] = r[pb - i]
    w[pb + i] = r[i]
    return w


def combine_channel_map(n_chan, combine):
    """jchan (0-based) for every chan of polyphase_synthesis.m:198-238."""
    jmap = np.arange(n_chan)
    if combine > 1:
        per_coarse = n_chan // combine
        for chan in range(n_chan):
            fine = (chan + per_coarse // 2) % n_chan                 # :218
            coarse = fine // per_coarse                              # :223
            fine = fine - coarse * per_coarse                        # :224
            coarse = (coarse + combine // 2) % combine               # :229
            fine = (fine + per_coarse // 2) % per_coarse             # :233
            jmap[chan] = coarse * per_coarse + fine                  # :234
    return jmap


# --------------------------------------------------------------------------
# inversion: polyphase_synthesis.m
# --------------------------------------------------------------------------
def _synth_params(n_chan, n_dat, nfft, os_factor, overlap):
    keep = nfft - 2 * overlap
    n_blocks = max((n_dat - 2 * overlap) // keep, 0) if keep > 0 else 0
    out_n = os_factor.normalize(nfft) * n_chan
    out_ov = os_factor.normalize(overlap) * n_chan
    if int(out_n) != out_n or int(out_ov) != out_ov:
        raise ValueError("input_fft_length and input_overlap must be divisible by os_factor.nu")
    out_n, out_ov = int(out_n), int(out_ov)
    fn_width = (nfft * os_factor.de) // os_factor.nu
    return keep, n_blocks, out_n, out_ov, out_n - 2 * out_ov, fn_width


def polyphase_synthesis_literal(x, spans_nyquist, input_fft_length, os_factor,
                                deripple=None, sample_offset=1, input_overlap=None,
                                temporal_taper=None, spectral_taper=None, combine=1):
    """matlab/polyphase_synthesis.m:99-316, loop for loop.

    x: (n_pol, n_chan, n_dat). ``deripple``: None or filter taps (apply).
    ``temporal_taper``: length-Nf real window or None; ``spectral_taper``:
    length n_chan*FN_width real window or None. sample_offset is 1-based.
    Scratch arrays are double (:152,:154); ``out`` takes the class of x (:131)."""
    os_factor = _os(os_factor)
    x = np.asarray(x)
    if input_overlap is None:
        input_overlap = input_fft_length // 8                      # :78
    x = x[:, :, sample_offset - 1:]                                # :99
    n_pol, n_chan, n_dat = x.shape
    cdt = _cplx_dtype(x.dtype)
    nfft = input_fft_length
    keep, n_blocks, out_n, out_ov, out_keep, fnw = _synth_params(n_chan, n_dat, nfft, os_factor, input_overlap)
    fnw2 = fnw // 2
    discard_2 = (nfft - fnw) // 2                                  # :136
    out = np.zeros((n_pol, 1, n_blocks * out_keep), dtype=cdt)     # :131
    if deripple is not None:
        pb = fnw // 2                                              # :142
        resp = deripple_response(deripple, n_chan, fnw)            # :143-149
    jmap = combine_channel_map(n_chan, combine)
    ffff = np.zeros(n_chan * fnw, dtype=np.complex128)             # :152
    for n in range(n_blocks):                                      # :163
        for i_pol in range(n_pol):
            s = keep * n                                           # :165
            in_dat = x[i_pol, :, s: s + nfft].astype(np.complex128)   # :171
            if temporal_taper is not None:
                in_dat = in_dat * np.asarray(temporal_taper)[None, :]  # :179
            spectra = np.fft.fft(in_dat.T, nfft, axis=0)           # :184-185 (Nf x n_chan)
            spectra = np.fft.fftshift(spectra, axes=0)             # :188
            fn = np.zeros((fnw, n_chan), dtype=cdt)                # :192
            for chan in range(n_chan):                             # :193
                fn[:, chan] = spectra[discard_2: discard_2 + fnw, jmap[chan]]   # :240
                if deripple is not None:
                    for ii in range(pb):                           # :244-250
                        fn[ii, chan] = fn[ii, chan] * resp[pb - ii]
                        fn[pb + ii, chan] = fn[pb + ii, chan] * resp[ii]
                if not spans_nyquist:
                    ffff[chan * fnw: (chan + 1) * fnw] = fn[:, chan]   # :254
            if spans_nyquist:                                      # :265-278
                ffff[:fnw2] = fn[fnw2:, 0]
                ffff[n_chan * fnw - fnw2:] = fn[:fnw2, 0]
                for chan in range(1, n_chan):
                    st = (chan - 1) * fnw + fnw2
                    ffff[st: st + fnw] = fn[:, chan]
            if spectral_taper is not None:
                ffff = ffff * np.asarray(spectral_taper)           # :282
            iffff = np.fft.ifft(ffff) / (os_factor.nu / os_factor.de)   # :285
            out[i_pol, 0, out_keep * n: out_keep * (n + 1)] = iffff[out_ov: out_n - out_ov]   # :302
    return out


def polyphase_synthesis(x, spans_nyquist, input_fft_length, os_factor,
                        deripple=None, sample_offset=1, input_overlap=None,
                        temporal_taper=None, spectral_taper=None, combine=1,
                        block_chunk=32):
    """Vectorised form of polyphase_synthesis.m (same maths, batched over
    overlap-save blocks).  Always computes in double like the Matlab scratch
    arrays and casts to the class of ``x`` on output."""
    os_factor = _os(os_factor)
    x = np.asarray(x)
    if input_overlap is None:
        input_overlap = input_fft_length // 8
    x = x[:, :, sample_offset - 1:]
    n_pol, n_chan, n_dat = x.shape
    cdt = _cplx_dtype(x.dtype)
    nfft = input_fft_length
    keep, n_blocks, out_n, out_ov, out_keep, fnw = _synth_params(n_chan, n_dat, nfft, os_factor, input_overlap)
    fnw2 = fnw // 2
    d2 = (nfft - fnw) // 2
    out = np.zeros((n_pol, 1, n_blocks * out_keep), dtype=cdt)
    w_fn = deripple_fn_weights(deripple, n_chan, fnw) if deripple is not None else None
    jmap = combine_channel_map(n_chan, combine)
    taper = None if temporal_taper is None else np.asarray(temporal_taper, dtype=np.float64)
    staper = None if spectral_taper is None else np.asarray(spectral_taper, dtype=np.float64)
    for p in range(n_pol):
        xp = np.ascontiguousarray(x[p])
        for n0 in range(0, n_blocks, block_chunk):
            nb = min(block_chunk, n_blocks - n0)
            # tiles[b, c, t]
            tiles = np.lib.stride_tricks.as_strided(
                xp[:, n0 * keep:], shape=(nb, n_chan, nfft),
                strides=(keep * xp.strides[1], xp.strides[0], xp.strides[1]), writeable=False)
            tiles = tiles[:, jmap, :].astype(np.complex128)
            if taper is not None:
                tiles = tiles * taper
            sp = np.fft.fftshift(np.fft.fft(tiles, axis=2), axes=2)[:, :, d2: d2 + fnw]
            if w_fn is not None:
                sp = sp * w_fn
            if spans_nyquist:
                flat = sp.reshape(nb, n_chan * fnw)
                ff = np.roll(flat, -fnw2, axis=1)
            else:
                ff = sp.reshape(nb, n_chan * fnw)
            if staper is not None:
                ff = ff * staper
            y = np.fft.ifft(ff, axis=1) * (os_factor.de / os_factor.nu)
            out[p, 0, n0 * out_keep: (n0 + nb) * out_keep] = y[:, out_ov: out_n - out_ov].reshape(-1)
    return out


# --------------------------------------------------------------------------
# streaming wrappers: FilterBank.m, InverseFilterBank.m, TwoStage*.m
# --------------------------------------------------------------------------
_ANALYSIS = {
    "polyphase_analysis": polyphase_analysis,
    "polyphase_analysis_padded": polyphase_analysis_padded,
}


def matlab_round(z):
    """Matlab `round` (FilterBank.m:82,112): ties go AWAY from zero, applied to the real and
    imaginary parts separately (numpy's `round` sends ties to the even integer)."""
    z = np.asarray(z)
    def r(a):
        return np.sign(a) * np.floor(np.abs(a) + 0.5)
    if np.iscomplexobj(z):
        return (r(z.real) + 1j * r(z.imag)).astype(z.dtype)
    return r(z).astype(z.dtype)


class FilterBank:
    """matlab/FilterBank.m:65-128 (tail buffering, trim to multiple of nu,
    optional integer rounding)."""

    def __init__(self, filt_coeff, n_chan, os_factor, analysis_function="polyphase_analysis",
                 rnd_input=False, rms_input=0.0, rnd_output=False, rms_output=0.0):
        self.filt_coeff = np.asarray(filt_coeff).ravel()
        self.n_chan = n_chan
        self.os_factor = _os(os_factor)
        self.analysis_function = analysis_function
        self.rnd_input, self.rms_input = rnd_input, rms_input
        self.rnd_output, self.rms_output = rnd_output, rms_output
        self.input_buffer = None
        self.buffered_samples = 0
        self._lowcbf_first = True

    def _analysis(self, x):
        if self.analysis_function == "polyphase_analysis_lowcbf":
            first, self._lowcbf_first = self._lowcbf_first, False
            return polyphase_analysis_lowcbf(x, self.filt_coeff, do_padding=first)
        return _ANALYSIS[self.analysis_function](x, self.filt_coeff, self.n_chan, self.os_factor)

    def execute(self, x):
        x = np.asarray(x)
        if self.rnd_input:                                          # :75-83
            scale = 1.0
            if self.rms_input > 0:
                scale = self.rms_input / np.sqrt(np.var(x, ddof=1))
            x = matlab_round(scale * x)
        if self.buffered_samples > 0:                               # :85-88
            x = np.concatenate([self.input_buffer, x], axis=2)
            self.buffered_samples = 0
        n_in = x.shape[2]
        out = self._analysis(x)                                     # :91
        rem = out.shape[2] % self.os_factor.nu                      # :93-104
        if rem:
            out = out[:, :, : out.shape[2] - rem]
        if self.rnd_output:                                         # :106-113
            scale = 1.0
            if self.rms_output:
                scale = self.rms_output / np.sqrt(np.var(out, ddof=1))
            out = matlab_round(scale * out)
        input_idat = out.shape[2] * self.n_chan * self.os_factor.de // self.os_factor.nu   # :119
        self.buffered_samples = n_in - input_idat
        if self.buffered_samples > 0:
            self.input_buffer = x[:, :, input_idat:]
        return out


class InverseFilterBank:
    """matlab/InverseFilterBank.m:63-137.  Note :90 forces deripple off."""

    def __init__(self, filt_coeff, n_fft, overlap, os_factor, temporal_taper="tukey",
                 critical=False, combine=1, sample_offset=0, spectral_taper=None):
        self.filt_coeff = filt_coeff
        self.n_fft, self.overlap = n_fft, overlap
        self.os_factor = _os(os_factor)
        self.temporal_taper = make_temporal_taper(temporal_taper, n_fft, overlap)
        self.spectral_taper_name = spectral_taper
        self.critical, self.combine, self.sample_offset = critical, combine, sample_offset
        self.input_buffer = None
        self.buffered_samples = 0

    def execute(self, x):
        x = np.asarray(x)
        if self.buffered_samples > 0:
            x = np.concatenate([self.input_buffer, x], axis=2)
            self.buffered_samples = 0
        n_pol, n_chan, n_dat = x.shape
        if not np.iscomplexobj(x):
            raise ValueError("polyphase_synthesis input data are real-valued!")   # :83-85
        staper = None
        if self.spectral_taper_name == "hann":
            fnw = (self.n_fft * self.os_factor.de) // self.os_factor.nu
            staper = spectral_hann(n_chan * fnw)
        out = polyphase_synthesis(x, not self.critical, self.n_fft, self.os_factor, None,
                                  self.sample_offset + 1, self.overlap, self.temporal_taper,
                                  staper, self.combine)
        nu, de = self.os_factor.nu, self.os_factor.de
        while True:                                                 # :104-124
            input_idat = out.shape[2] * nu // (n_chan * de)
            self.buffered_samples = n_dat - input_idat
            rem = self.buffered_samples % nu
            if rem == 0:
                break
            self.buffered_samples += nu - rem
            input_idat = n_dat - self.buffered_samples
            out = out[:, :, : input_idat * (n_chan * de) // nu]
        if self.buffered_samples > 0:
            self.input_buffer = x[:, :, input_idat:]
        return out


class TwoStageFilterBank:
    """matlab/TwoStageFilterBank.m:58-116 (stage 1 on pol 1 only, one stage-2
    FilterBank per coarse channel, optional critical chomp, single output)."""

    def __init__(self, cfg1: dict, cfg2: dict | None = None, critical=False, single=False):
        cfg2 = cfg2 or cfg1
        self.cfg2 = cfg2
        self.stage1 = FilterBank(**cfg1)
        self.stage2 = [FilterBank(**cfg2) for _ in range(self.stage1.n_chan)]
        self.critical, self.single = critical, single

    def execute(self, x):
        out1 = self.stage1.execute(x)                               # :69
        os_ = self.stage1.os_factor
        nch1 = self.stage1.n_chan
        nch2_orig = self.stage2[0].n_chan
        nch2 = nch2_orig * os_.de // os_.nu if self.critical else nch2_orig   # :81-83
        offset = nch2_orig - nch2
        if self.single:
            nch1 = 1
        out = None
        for ich in range(nch1):                                     # :92
            tmp = self.stage2[ich].execute(out1[0:1, ich:ich + 1, :])
            if out is None:
                out = np.zeros((1, nch1 * nch2, tmp.shape[2]), dtype=np.complex64)   # :99 'single'
            h = nch2 // 2
            out[0, ich * nch2: ich * nch2 + h, :] = tmp[0, :h, :]                     # :104
            # :105  (nch2/2:nch2) 1-based  ->  0-based h-1 .. nch2-1  (index h-1 is written twice)
            out[0, ich * nch2 + h - 1: (ich + 1) * nch2, :] = tmp[0, h - 1 + offset: nch2 + offset, :]
        return out


class TwoStageInverseFilterBank:
    """matlab/TwoStageInverseFilterBank.m:75-157 (inverts stage 2 only)."""

    def __init__(self, cfg2: dict, nch2: int, combine=1, single=False):
        self.cfg2 = cfg2
        self.nch2 = nch2          # fine channels per coarse channel present in the input
        self.combine, self.single = combine, single
        self.stage2 = {}

    def execute(self, x, stage2_nchan):
        x = np.asarray(x)
        os_ = _os(self.cfg2["os_factor"])
        nchan = x.shape[1]
        nch_out = nchan // self.nch2
        crit_n = os_.normalize(stage2_nchan)
        if self.nch2 == crit_n:
            critical = True
        elif self.nch2 == stage2_nchan:
            critical = False
            if self.combine > 1:
                raise ValueError("cannot combine oversampled coarse channels")   # :110-112
        else:
            raise ValueError("invalid nchan")                                    # :113-115
        nch_in = self.nch2 * self.combine
        nch_out //= self.combine
        if self.single:
            nch_out = 1
        out = None
        for ich in range(nch_out):
            if ich not in self.stage2:
                self.stage2[ich] = InverseFilterBank(**self.cfg2)
            ifb = self.stage2[ich]
            ifb.critical, ifb.combine = critical, self.combine
            tmp = ifb.execute(x[0:1, ich * nch_in: (ich + 1) * nch_in, :])
            if out is None:
                out = np.zeros((1, nch_out, tmp.shape[2]), dtype=np.complex128)
            out[0, ich, :] = tmp[0, 0, :]
        return out


# --------------------------------------------------------------------------
# test-vector generators
# --------------------------------------------------------------------------
def complex_sinusoid(n, freqs, phases, bin_offset=0.0, dtype=np.complex64):
    """python/data_gen/generate_test_vector.py:24-48."""
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


def time_domain_impulse(n, offsets, widths, dtype=np.complex64):
    """python/data_gen/generate_test_vector.py:51-71."""
    if not hasattr(offsets, "__iter__"):
        offsets, widths = [offsets], [widths]
    sig = np.zeros(n, dtype=dtype)
    for i in range(len(offsets)):
        offset = offsets[i]
        if offset < 1.0:
            offset = int(offsets[i] * n)
        sig[int(offset): int(offset) + int(widths[i])] = 1.0
    return sig


def impulse(nsample, offset, amplitude=1.0, current=0, dtype=np.complex64):
    """matlab/Impulse.m:13-41 without the 1e-6 noise."""
    x = np.zeros(nsample, dtype=dtype)
    off = offset - current
    if 0 <= off < nsample:
        x[off] = amplitude
    return x


def pure_tone(nsample, frequency, amplitude=1.0, current=0, dtype=np.complex64):
    """matlab/PureTone.m:22-26: amplitude*exp(j 2 pi f (t+current))."""
    t = np.arange(nsample, dtype=np.float64) + current
    return (amplitude * np.exp(2j * np.pi * frequency * t)).astype(dtype)


def frequency_comb_params(n_chan=1, two_stage=False, invert=False, comb="", nharmonic=32):
    """matlab/sgcht.m:394-432 amplitudes and frequencies of the comb."""
    amplitudes = np.linspace(1.0, math.sqrt(2.0), nharmonic)
    fmin = -0.5 + 1.0 / (nharmonic * 4)
    fmax = fmin + (nharmonic - 1.0) / nharmonic
    if comb == "coarse":
        fmin, fmax = fmin / n_chan, fmax / n_chan
    elif comb == "fine":
        fmin, fmax = fmin / n_chan ** 2, fmax / n_chan ** 2
    elif n_chan > 1:
        nch = n_chan ** 2 if two_stage else n_chan
        if invert:
            nch = nch // n_chan
        if nch > 1:
            fmin += 1.0 / (nch * 4)
            fmax += 1.0 / (nch * 4)
    return amplitudes, np.linspace(fmin, fmax, nharmonic)


def frequency_comb(nsample, amplitudes, frequencies, current=0, dtype=np.complex64):
    """matlab/FrequencyComb.m:36-46: sum of PureTone generators (each tone is
    rounded to single before the sum, as the Matlab does)."""
    x = np.zeros(nsample, dtype=dtype)
    for a, f in zip(amplitudes, frequencies):
        x = x + pure_tone(nsample, f, a, current, dtype)
    return x


# --------------------------------------------------------------------------
# prototype filter design (the .mat files named by the configs are git-ignored
# upstream; re-created with scipy.signal.firls)
# --------------------------------------------------------------------------
def design_pfb_fir_filter(n_chan, os_factor, n_taps_per_chan):
    """matlab/design_PFB_FIR_filter.m:24-48: firls order n_chan*taps_per_chan,
    bands [0,Fp,Fs,1], Fp=1/n_chan, Fs=(2*OS-1)/n_chan, weights [1,15]."""
    from scipy.signal import firls
    os_factor = _os(os_factor)
    os_f = os_factor.nu / os_factor.de
    if os_f == 1:
        os_f += 0.1
    fp = 1.0 / n_chan
    fs = (2 * os_f - 1) / n_chan
    order = n_chan * n_taps_per_chan
    return firls(order + 1, [0, fp, fs, 1.0], [1, 1, 0, 0], weight=[1, 15])


def design_pfb_fir_filter_two_stage(n_chan, os_factor, os_taps_per_chan=28, zero_stuff_factor=32):
    """matlab/design_PFB_FIR_filter_two_stage.m:42-78."""
    from scipy.signal import firls
    os_factor = _os(os_factor)
    os_f = os_factor.nu / os_factor.de
    n_taps = int(round(os_taps_per_chan * n_chan / os_f))
    n1 = n_taps // zero_stuff_factor
    fp = 1.0 / n_chan
    fs = (2 * os_f - 1) / n_chan
    h0 = firls(n1 + 1, [0, fp * zero_stuff_factor, 0.998 * fs * zero_stuff_factor, 1.0],
               [1, 1, 0, 0], weight=[1, 15])
    h1 = np.fft.fft(np.fft.ifftshift(h0))
    hz = np.concatenate([h1[: n1 // 2 + 1], np.zeros(n1 * (zero_stuff_factor - 1)), h1[n1 // 2 + 1:]])
    return np.real(np.fft.fftshift(np.fft.ifft(hz)))
