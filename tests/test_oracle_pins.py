"""Pins the oracle on the known-answer / behavioural criteria the reference's own tests
hold for this path - the reference ships no golden vectors."""
import numpy as np
import pytest

from oracle import pfb_oracle as o


@pytest.fixture(scope="module")
def h(pst_taps):
    return pst_taps / pst_taps.sum()


def _db(a):
    return 20 * np.log10(np.abs(a) + 1e-300)


def test_sizes_known_answers():
    # calc_output_nbins.m:17-27 evaluated by hand for the Low config
    assert o.calc_output_nbins(147456, 256, "4/3", 3072, 256, 48) == 61440
    # python/test/test_dspsr_util.py:77-81: output_fft_length 229376 = 8 * 32768 * 7/8
    assert o.normalize("8/7", 32768) * 8 == 229376
    assert o.normalize({"nu": 8, "de": 7}, 32) == 28 and o.multiply("8/7", 28) == 32        # normalize.m / multiply.m examples
    assert o.pad_filter(np.ones(81), 8).size == 88                                             # pad_filter.m
    # SURVEY §8a shapes: cfg1 (1,8,37010), cfg2 nblocks 87365
    assert (2 ** 18 - 3072) // 7 == 37010 and (2 ** 24 - 3072) // 192 == 87365


@pytest.mark.parametrize("offset", [20000, 30000, 30000 + 30720 - 5])
def test_impulse_round_trip_low(h, offset):
    """TestImpulse.m:46-73 with sgcht.m:466-473: peak at offset + fir_offset - (output_overlap - 1 + kludge),
    (dir, kludge) = (-1, 1) for polyphase_analysis; all other samples below -60 dB."""
    n = 256 * 192 * 3
    x = np.zeros((1, 1, n), dtype=np.complex64)
    x[0, 0, offset] = 1
    ch = o.polyphase_analysis(x, h, 256, "4/3")
    y = o.polyphase_synthesis(ch, True, 256, "4/3", h, 1, 48, o.window_tukey(256, 48))[0, 0]
    expect = offset + (-1) * (3072 // 2) - (36 * 256 - 1 + 1)
    assert int(np.argmax(np.abs(y))) == expect
    rest = np.abs(y).copy()
    rest[expect - 1: expect + 2] = 0
    assert _db(rest.max()) < -60


def test_impulse_round_trip_lowcbf(pst_taps):
    """config 'lowpsi': polyphase_analysis_lowcbf, (dir, kludge) = (+1, 193), 216 kept channels."""
    n = 216 * 192 * 3 + 3072
    offset = 20000
    x = np.zeros((1, 1, n), dtype=np.complex128)
    x[0, 0, offset] = 1
    ch = o.polyphase_analysis_lowcbf(x, pst_taps)
    y = o.polyphase_synthesis(ch, True, 256, "4/3", None, 1, 48, o.window_tukey(256, 48))[0, 0]
    # PSTFilterbank.m:3-4: the 1536-sample pre-pad makes output sample k centred on input sample 192 k;
    # 216 kept channels -> 162 output samples per fine sample, 48 fine samples of overlap dropped
    expect = round((offset / 192 - 48) * 162)
    assert abs(int(np.argmax(np.abs(y))) - expect) <= 1
    # the robust pin: the inverted impulse is a single dominant peak, leakage < -60 dB beyond +-2 samples
    pk = int(np.argmax(np.abs(y)))
    rest = np.abs(y).copy() / np.abs(y[pk])
    rest[pk - 2: pk + 3] = 0
    assert _db(rest.max()) < -60


def test_pure_tone_round_trip_low(h):
    """TestPureTone.m:53-89: FFT peak at f*nfft, every other bin below -60 dB."""
    n = 256 * 192 * 4
    nfft_out = 30720 * 2
    f = 300 / nfft_out                      # an exact bin of the tested length
    x = o.pure_tone(n, f, dtype=np.complex128)[None, None, :]
    ch = o.polyphase_analysis(x, h, 256, "4/3")
    y = o.polyphase_synthesis(ch, True, 256, "4/3", h, 1, 48, o.window_tukey(256, 48))[0, 0]
    y = y[:nfft_out]
    sp = _db(np.fft.fft(y) / nfft_out)
    pk = int(np.argmax(sp))
    assert pk == round(f * nfft_out)
    sp -= sp[pk]
    sp[pk] = -999
    assert sp.max() < -60


def _comb_test(x_out, freqs, os_factor, two_stage=False, invert=False, critical=False, max_nfft=8192):
    """TestFrequencyComb.m:80-112 (returns 0 on pass)."""
    npol, nchan, _ = x_out.shape
    for ichan in range(nchan):
        x = x_out[0, ichan, :]
        nfft = min(x.size, max_nfft)
        fft_in = np.abs(np.fft.fft(x[:nfft]) / (nfft * nchan))
        level = 2 if two_stage else (1 if nchan > 1 else 0)
        if invert:
            level -= 1
        if critical:
            level -= 1
        hfac = nchan * nfft
        for _ in range(max(level, 0)):
            hfac = o.normalize(os_factor, hfac)
        for fr in freqs:
            jchan = int(np.floor(fr * nchan)) % nchan
            if jchan == ichan:
                offset = ichan / nchan
                iharm = int(np.floor((fr - offset) * hfac)) % nfft
                if fft_in[iharm] < 0.5:
                    return -1
    return 0


def test_frequency_comb_none_and_channelised(h):
    """test_sgcht.m options 1-2: the comb itself, and the comb after the Low channeliser."""
    amps, freqs = o.frequency_comb_params()
    x = o.frequency_comb(8192, amps, freqs, dtype=np.complex128)[None, None, :]
    assert _comb_test(x, freqs, "1/1") == 0
    amps, freqs = o.frequency_comb_params(n_chan=256)
    n = 3072 + 192 * 2100
    x = o.frequency_comb(n, amps, freqs, dtype=np.complex128)[None, None, :]
    fb = o.FilterBank(h, 256, "4/3")
    ch = fb.execute(x)
    assert ch.shape[2] % 4 == 0
    assert _comb_test(ch[:, :, :2048], freqs, "4/3") == 0


def test_frequency_comb_inverted(h):
    """test_sgcht.m option 3: channelise + invert (InverseFilterBank forces deripple off)."""
    amps, freqs = o.frequency_comb_params(n_chan=256, invert=True)
    n = 3072 + 192 * (160 * 3 + 96 + 8)
    x = o.frequency_comb(n, amps, freqs, dtype=np.complex128)[None, None, :]
    ch = o.FilterBank(h, 256, "4/3").execute(x)
    y = o.InverseFilterBank(h, 256, 48, "4/3", "tukey").execute(ch)
    assert y.shape[2] >= 8192
    assert _comb_test(y, freqs, "4/3", invert=True) == 0


def test_chop_alignment_rule(h):
    """chop.m:36-37 / purity.py:86-88: input[to_chomp:] lines up with the inverted series."""
    rng = np.random.default_rng(2)
    n = 256 * 192 * 3
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * np.hanning(n)
    # band-limit so the PFB passband ripple is the only error
    x = np.convolve(x, np.ones(8) / 8, mode="same")[None, None, :]
    ch = o.polyphase_analysis(x, h, 256, "4/3")
    y = o.polyphase_synthesis(ch, True, 256, "4/3", h, 1, 48, o.window_tukey(256, 48))[0, 0]
    to_chomp = (36 * 256 - 1 + 1) - (-1) * (3072 // 2)
    a = x[0, 0, to_chomp: to_chomp + y.size]
    assert np.abs(y - a).max() / np.abs(a).max() < 2e-3
    # and the python driver's shift differs by exactly one sample ((taps-1)//2 vs floor(L/2))
    assert 36 * 256 + (3072 - 1) // 2 == to_chomp - 1


def test_two_stage_shapes_and_critical_chomp():
    """TwoStageFilterBank.m:58-116: stage-2 per coarse channel, critical keeps nch2*de/nu fine channels."""
    rng = np.random.default_rng(3)
    h8 = o.design_pfb_fir_filter(8, "4/3", 6)
    cfg = dict(filt_coeff=h8, n_chan=8, os_factor="4/3")
    x = (rng.standard_normal((1, 1, 4000)) + 1j * rng.standard_normal((1, 1, 4000)))
    full = o.TwoStageFilterBank(cfg).execute(x)
    crit = o.TwoStageFilterBank(cfg, critical=True).execute(x)
    assert full.shape[1] == 64 and crit.shape[1] == 8 * 6 and full.dtype == np.complex64
    # TwoStageFilterBank.m:104-105: the two slice assignments overlap at index nch2/2,
    # so fine channel 2 is overwritten by channel 4 of the oversampled set
    assert np.array_equal(crit[0, :2], full[0, :2])
    assert np.array_equal(crit[0, 2:6], full[0, 4:8])
    inv = o.TwoStageInverseFilterBank(dict(filt_coeff=h8, n_fft=24, overlap=3, os_factor="4/3"), nch2=6, combine=2)
    y = inv.execute(crit, stage2_nchan=8)
    assert y.shape[1] == 8 // 2 and y.shape[2] > 0
