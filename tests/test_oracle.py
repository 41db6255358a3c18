"""Oracle self-consistency: literal Matlab transliteration == closed form."""
import numpy as np
import pytest

from oracle import pfb_oracle as o


def nerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def cnoise(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def test_bunton_literal_vs_closed_low(pst_taps):
    rng = np.random.default_rng(0)
    h = pst_taps / pst_taps.sum()
    x = cnoise(rng, (2, 1, 3072 + 192 * 24 + 5))
    a = o.polyphase_analysis_literal(x, h, 256, "4/3")
    b = o.polyphase_analysis(x, h, 256, "4/3")
    assert a.shape == (2, 256, 24)
    assert nerr(b, a) < 1e-12


@pytest.mark.parametrize("nchan,osf,ntaps", [(8, "8/7", 3072), (8, "8/7", 81), (16, "4/3", 97), (32, "32/27", 32 * 5), (8, "1/1", 64)])
def test_bunton_literal_vs_closed_small(nchan, osf, ntaps):
    rng = np.random.default_rng(1)
    h = rng.standard_normal(ntaps)
    x = cnoise(rng, (1, 1, ntaps + 700))
    a = o.polyphase_analysis_literal(x, h, nchan, osf)
    b = o.polyphase_analysis(x, h, nchan, osf)
    assert a.shape[2] > 0 and nerr(b, a) < 1e-12


def test_bunton_single_precision_class():
    rng = np.random.default_rng(2)
    x = cnoise(rng, (1, 1, 2000)).astype(np.complex64)
    a = o.polyphase_analysis_literal(x, rng.standard_normal(81), 8, "8/7")
    b = o.polyphase_analysis(x, np.zeros(81) + 1, 8, "8/7")
    assert a.dtype == np.complex64 and b.dtype == np.complex64


@pytest.mark.parametrize("nchan,osf,ntaps", [(16, "8/7", 16 * 25 + 1), (8, "4/3", 50), (32, "8/7", 32 * 3)])
def test_padded_literal_vs_closed(nchan, osf, ntaps):
    rng = np.random.default_rng(3)
    h = rng.standard_normal(ntaps)
    x = cnoise(rng, (2, 1, 3000))
    a = o.polyphase_analysis_padded_literal(x, h, nchan, osf)
    b = o.polyphase_analysis_padded(x, h, nchan, osf)
    assert nerr(b, a) < 1e-12
    # first (pre-shift) output block is identically zero: it
    # lands at index nblocks - shift after the circular delay compensation
    os_ = o.OSFactor.from_str(osf)
    step = nchan * os_.de // os_.nu
    sds = int(np.ceil((ntaps - 1) / 2 / step))
    assert np.abs(a[:, :, a.shape[2] - sds]).max() == 0.0


@pytest.mark.parametrize("pad", [True, False])
def test_lowcbf_literal_vs_closed(pst_taps, pad):
    rng = np.random.default_rng(4)
    x = cnoise(rng, (2, 1, 3072 + 192 * 5))
    a = o.polyphase_analysis_lowcbf(x, pst_taps, literal=True, do_padding=pad)
    b = o.polyphase_analysis_lowcbf(x, pst_taps, do_padding=pad)
    assert a.shape[1] == 216 and a.shape[2] == (13 if pad else 5)
    assert nerr(b, a) < 1e-12


CASES = [
    (8, "8/7", 64, 8, True, 1, True, "tukey", False),
    (8, "8/7", 64, 8, False, 1, False, None, False),
    (16, "4/3", 48, 12, False, 4, False, "tukey", False),
    (8, "4/3", 24, 6, True, 1, True, "top_hat", True),
    (12, "4/3", 24, 0, False, 2, True, "no_window", True),
]


@pytest.mark.parametrize("nch,osf,nfft,ov,spans,comb,dr,tap,st", CASES)
def test_synthesis_literal_vs_vectorised(nch, osf, nfft, ov, spans, comb, dr, tap, st):
    rng = np.random.default_rng(5)
    os_ = o.OSFactor.from_str(osf)
    x = cnoise(rng, (2, nch, nfft * 5 + 7))
    hh = rng.standard_normal(nch * 6 + 1)
    tt = o.make_temporal_taper(tap, nfft, ov)
    fnw = nfft * os_.de // os_.nu
    stp = o.spectral_hann(nch * fnw) if st else None
    a = o.polyphase_synthesis_literal(x, spans, nfft, osf, hh if dr else None, 2, ov, tt, stp, comb)
    b = o.polyphase_synthesis(x, spans, nfft, osf, hh if dr else None, 2, ov, tt, stp, comb, block_chunk=3)
    assert a.shape[2] > 0 and nerr(b, a) < 1e-12


def test_combine_map_is_permutation():
    for n, c in [(16, 4), (192 * 16, 16), (8, 2)]:
        m = o.combine_channel_map(n, c)
        assert sorted(m.tolist()) == list(range(n))
    assert (o.combine_channel_map(8, 1) == np.arange(8)).all()


def test_windows():
    w = o.window_tukey(256, 48)
    assert w[0] == 0 and w[-1] == 0 and (w[48:208] == 1).all()
    assert np.allclose(w[:48], o.hann(96)[:48]) and np.allclose(w[208:], o.hann(96)[48:])
    t = o.window_top_hat(16, 3)
    assert t[:3].sum() == 0 and t[-3:].sum() == 0 and t[3:13].sum() == 10
    s = o.spectral_hann(12)
    assert np.allclose(s, np.roll(o.hann(12), 6))


def test_freqz_matches_scipy(pst_taps):
    from scipy.signal import freqz
    h = pst_taps / pst_taps.sum()
    w, hh = freqz(h, 1, worN=256 * 96)
    mine = o.freqz_half(h, 256 * 96, 97)
    assert np.abs(mine - hh[:97]).max() < 1e-12
    r = o.deripple_response(h, 256, 192)
    assert r.shape == (97,) and abs(r[0] - 1.0) < 1e-3


def test_streaming_filterbank_equals_oneshot(pst_taps):
    """FilterBank.m tail carry: blocks fed in chunks == one-shot (trimmed)."""
    rng = np.random.default_rng(6)
    h = pst_taps / pst_taps.sum()
    x = cnoise(rng, (1, 1, 3072 + 192 * 37))
    fb = o.FilterBank(h, 256, "4/3")
    outs = [fb.execute(x[:, :, :5000]), fb.execute(x[:, :, 5000:8000]), fb.execute(x[:, :, 8000:])]
    got = np.concatenate(outs, axis=2)
    ref = o.polyphase_analysis(x, h, 256, "4/3")
    assert got.shape[2] % 4 == 0 and got.shape[2] <= ref.shape[2]
    assert nerr(got, ref[:, :, : got.shape[2]]) < 1e-12


def test_streaming_inverse_equals_oneshot():
    rng = np.random.default_rng(7)
    x = cnoise(rng, (1, 8, 64 * 9 + 5))
    ifb = o.InverseFilterBank(None, 64, 8, "8/7", "tukey")
    # chunk lengths are multiples of nu so InverseFilterBank.m:113-123 never has
    # to truncate (a truncation re-phases the overlap-save grid upstream too)
    outs = [ifb.execute(x[:, :, :200]), ifb.execute(x[:, :, 200:432]), ifb.execute(x[:, :, 432:])]
    got = np.concatenate(outs, axis=2)
    ref = o.polyphase_synthesis(x, True, 64, "8/7", None, 1, 8, o.window_tukey(64, 8))
    assert got.shape[2] > 0 and nerr(got, ref[:, :, : got.shape[2]]) < 1e-12


def test_generators():
    s = o.complex_sinusoid(1000, [0.01], [np.pi / 4], 0.0)
    assert s.dtype == np.complex64 and abs(np.fft.fft(s)[10]) > 999
    i = o.time_domain_impulse(1000, [0.11], [1])
    assert i[110] == 1 and i.sum() == 1
    a, f = o.frequency_comb_params()
    assert len(a) == 32 and abs(f[0] - (-0.5 + 1 / 128)) < 1e-15 and abs(a[-1] - 2 ** 0.5) < 1e-15
    c = o.frequency_comb(256, a, f)
    assert c.shape == (256,)


def test_matlab_round_ties_away_from_zero():
    """FilterBank.m:82,112 use Matlab round(): 0.5 -> 1, -0.5 -> -1, 2.5 -> 3 (numpy would give 0, -0, 2)."""
    z = np.array([0.5, -0.5, 1.5, 2.5, -2.5, 0.49, -3.51])
    assert np.array_equal(o.matlab_round(z), [1, -1, 2, 3, -3, 0, -4])
    zc = np.array([2.5 - 0.5j, -1.5 + 4.5j])
    assert np.array_equal(o.matlab_round(zc), [3 - 1j, -2 + 5j])
    h = np.ones(16) / 16
    x = (np.arange(400) % 7 - 3 + 0.5).astype(np.complex128)[None, None, :]
    a = o.FilterBank(h, 8, "8/7", rnd_input=True).execute(x)
    b = o.FilterBank(h, 8, "8/7").execute(o.matlab_round(x))
    assert np.array_equal(a, b)
