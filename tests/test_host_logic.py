"""CPU tests of the host-side mirror of the reference interface (no GPU, no compute)."""
import os
import re

import numpy as np
import pytest

from oracle import pfb_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol():
    """The shared library loads and exports exactly what include/pfb_b200.h declares."""
    from ska_pst_dsp_model_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "pfb_b200.h")).read()
    declared = set(re.findall(r"\b(pfb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib.lib, name), f"{name} declared in pfb_b200.h but not exported"
    assert declared == set(lib.EXPORTS)
    assert lib.lib.pfb_version() >= 100


def test_plan_argument_errors_need_no_gpu_compute():
    import ctypes
    from ska_pst_dsp_model_b200 import lib
    h = ctypes.c_void_p()
    taps = (ctypes.c_double * 4)(1, 2, 3, 4)
    rc = lib.lib.pfb_analysis_plan_create(ctypes.byref(h), 0, 8, 7, 8, taps, 4, 0, 0)     # os_de > os_nu
    assert rc == -1 and b"bad argument" in lib.lib.pfb_last_error()
    rc = lib.lib.pfb_analysis_plan_create(ctypes.byref(h), 0, 8, 8, 7, taps, 4, 5, 0)     # bad dtype
    assert rc == -1
    rc = lib.lib.pfb_inverse_plan_create(ctypes.byref(h), 8, 8, 7, 60, 8, None, None, 0, 1, 1, 0, 0)
    assert rc == -1 and b"divisible" in lib.lib.pfb_last_error()


def test_rational():
    from ska_pst_dsp_model_b200.rational import Rational
    r = Rational.from_str("8/7")
    assert r.normalize(32) == 28 and r.multiply(28) == 32 and str(r) == "8/7"
    assert Rational.from_str("4/3") == "4/3" and float(Rational(32, 27)) == 32 / 27
    assert Rational.from_str(r) is r
    # test_dspsr_util.py:77-81 pins output_fft_length = 8 ch * 32768 * 7/8
    assert r.normalize(32768) * 8 == 229376


def test_windows_match_oracle():
    from ska_pst_dsp_model_b200 import fft_windows as w
    assert np.array_equal(w.tukey_window(256, 48), o.window_tukey(256, 48))
    assert np.array_equal(w.top_hat_window(64, 8), o.window_top_hat(64, 8))
    assert np.array_equal(w.no_window(16), np.ones(16))
    assert np.allclose(w.hann_window(33), o.hann(33))
    with pytest.raises(KeyError):
        w.make_window("blackman", 8, 1)


def test_partialize_semantics():
    from ska_pst_dsp_model_b200.partialize import partialize

    @partialize
    def f(a, b=1, backend="x"):
        return (a, b, backend)

    g = f(backend="python")
    assert callable(g) and g(3) == (3, 1, "python") and g(3, b=2) == (3, 2, "python")
    assert f(5) == (5, 1, "x")


def test_dada_header_roundtrip_and_growth(tmp_path):
    from ska_pst_dsp_model_b200.formats import DADAFile, format_header, parse_header
    rng = np.random.default_rng(0)
    d = (rng.standard_normal((100, 4, 2)) + 1j * rng.standard_normal((100, 4, 2))).astype(np.complex64)
    f = DADAFile(str(tmp_path / "a.dump"))
    f.data = d
    f["OS_FACTOR"] = "8/7"
    f.dump_data()
    g = DADAFile(f.file_path).load_data()
    assert g["HDR_SIZE"] == "4096" and g["NCHAN"] == "4" and g["NPOL"] == "2" and g["NBIT"] == "32"      # test_read_dada_file.m:11-24
    assert np.array_equal(g.data, d) and g.nchan == 4 and g.npol == 2
    raw = open(f.file_path, "rb").read()
    assert raw.startswith(b"HDR_SIZE 4096\n") and raw[4095:4096] == b"\0" and len(raw) == 4096 + d.nbytes
    # payload order: pol fastest, then channel, then time, (re, im) interleaved
    pay = np.frombuffer(raw[4096:], dtype=np.float32)
    assert pay[0] == d[0, 0, 0].real and pay[1] == d[0, 0, 0].imag and pay[2] == d[0, 0, 1].real and pay[4] == d[0, 1, 0].real
    # header growth by doubling (write_header.m:22-26; test_write_dada_file.m:51,57 expects 8192)
    g["COEFF_0"] = ",".join("%0.6E" % c for c in np.arange(400))
    g.file_path = str(tmp_path / "b.dump")
    g.dump_data()
    h = DADAFile(g.file_path).load_data()
    assert h["HDR_SIZE"] == "8192" and np.array_equal(h.data, d) and len(h["COEFF_0"].split(",")) == 400
    # comments are skipped (read_header.m:20-22)
    assert "X" not in parse_header(b"HDR_SIZE 4096\n# X 1\nA b\n\0\0")
    f64 = DADAFile(str(tmp_path / "c.dump"))
    f64.data = d.astype(np.complex128)
    assert f64["NBIT"] == "64"
    f64.dump_data()
    assert np.array_equal(DADAFile(f64.file_path).load_data().data, d.astype(np.complex128))


def test_generate_test_vector_naming_and_content(tmp_path):
    from ska_pst_dsp_model_b200 import data_gen
    gen = data_gen.generate_test_vector(backend="python", n_bins=1000)
    f = gen([0.1], [1], domain_name="time", output_dir=str(tmp_path), n_pol=2)
    # python/test/test_data_gen.py:89-93
    assert os.path.basename(f.file_path) == "time_domain_impulse.1000.0.100-1.000.2.single.python.dump"
    assert f.data.shape == (1000, 1, 2) and f.data[100, 0, 0] == 1 and f.data.sum() == 2
    s = data_gen.generate_test_vector([0.01], [np.pi / 4], 0.0, n_bins=500, domain_name="freq", backend="python",
                                      output_dir=str(tmp_path))
    assert np.allclose(s.data[:, 0, 0], o.complex_sinusoid(500, [0.01], [np.pi / 4], 0.0))
    assert np.array_equal(data_gen.time_domain_impulse(50, [7], [2]), o.time_domain_impulse(50, [7], [2]))


def test_config_and_filter_resolution():
    from ska_pst_dsp_model_b200.data_gen.config import load_config
    from ska_pst_dsp_model_b200.filters import load_pst_taps, resolve_fir_filter
    low = load_config("low")
    assert low["channels"] == 256 and low["os_factor"] == "4/3" and low["input_overlap"] == 48
    mid = load_config("mid")
    assert mid["analysis_function"] == "polyphase_analysis_padded" and mid["fir_filter_taps"] == 100353
    t = load_pst_taps()
    # shipped taps invariants: 3072 integer taps, sum 16 777 241, peak at 1536, symmetric about it
    assert t.size == 3072 and t.sum() == 16777241 and t.argmax() == 1536 and np.array_equal(t[1:], t[1:][::-1])
    assert np.array_equal(resolve_fir_filter("whatever/PST_filtertaps.txt"), t)
    with pytest.raises(FileNotFoundError):
        resolve_fir_filter("missing.mat")


def test_designed_filters_match_oracle_designs():
    from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter
    h = design_pfb_fir_filter(256, 4, 3, 12)
    ho = o.design_pfb_fir_filter(256, "4/3", 12)
    assert h.size == 3073 and np.allclose(h, ho) and np.allclose(h, h[::-1]) and abs(h.sum() - 1) < 1e-3


def test_time_shard_plan_covers_stream():
    from ska_pst_dsp_model_b200.sharding import plan_time_shards
    n = 1 << 22
    for world in (1, 2, 3, 8):
        sh = plan_time_shards(n, world, 256, "4/3", 3072, 256, 48)
        k_total = (n - 3072) // 192
        nb_total = (k_total - 96) // 160
        assert sum(s.inv_nblocks for s in sh) == nb_total
        assert sh[0].own0 == 0 and sh[-1].own0 + sh[-1].own_n == n
        for a, b in zip(sh[:-1], sh[1:]):
            assert a.own0 + a.own_n == b.own0 and a.out0 + a.out_n == b.out0
            assert a.halo_n == 96 * 192 + 3072 - 192 and b.send_n == a.halo_n      # 2*ov*S + fl - S
            assert a.fine0 % 4 == 0 and b.own0 % 256 == 0                           # phase index restarts cleanly
        assert all(s.need0 + s.need_n <= n for s in sh)
    with pytest.raises(ValueError):
        plan_time_shards(3072 + 192 * 300, 8, 256, "4/3", 3072, 256, 48)


def test_purity_driver_parameters():
    """purity.py:79-99 bookkeeping (no compute)."""
    from ska_pst_dsp_model_b200.verify.purity import TestPurity
    t = TestPurity(n_test=5, os_factor="4/3", input_fft_length=256, input_overlap=48, fft_window="tukey",
                   deripple=True, channels=256, fir_filter_taps=3072, fir_filter_coeff_file_path="PST_filtertaps.txt",
                   blocks=3)
    assert t.block_size == 49152 and t.n_samples == 147456 and t.fft_size == 98304
    assert t.total_sample_shift == 36 * 256 + 1535
    assert list(t.time_domain_args["offset"]) == list(np.linspace(1, 147456, 5).astype(int))
    assert list(t.freq_domain_args["frequency"]) == list(np.linspace(1, 49152, 5).astype(int) * 3)


def test_host_round_trip_ramped_segments_tile_the_stream():
    """streaming.HostRoundTrip._ramp merges the fine time cuts into segments that ramp up at the head and down at
    the tail; they must tile inputs, fine blocks and outputs exactly like the cuts they were merged from."""
    from ska_pst_dsp_model_b200.sharding import plan_time_shards
    from ska_pst_dsp_model_b200.streaming import HostRoundTrip
    n = 1 << 22
    cuts = plan_time_shards(n, 32, 256, "4/3", 3072, 256, 48)
    for flat in (False, True):
        segs = HostRoundTrip._ramp(cuts, 8, flat=flat)
        assert segs[0].out0 == 0 and segs[0].need0 == cuts[0].need0 and segs[0].fine0 == cuts[0].fine0
        for a, b in zip(segs, segs[1:]):
            assert a.out0 + a.out_n == b.out0
            assert b.need0 < a.need0 + a.need_n              # the FIR history / overlap-save overlap is re-read
        last = segs[-1]
        assert last.out0 + last.out_n == cuts[-1].out0 + cuts[-1].out_n
        assert last.need0 + last.need_n == cuts[-1].need0 + cuts[-1].need_n
        assert last.fine0 + last.fine_n == cuts[-1].fine0 + cuts[-1].fine_n
        assert sum(s.out_n for s in segs) == sum(c.out_n for c in cuts)
    ramped = HostRoundTrip._ramp(cuts, 8)
    assert ramped[0].out_n < ramped[len(ramped) // 2].out_n and ramped[-1].out_n < ramped[len(ramped) // 2].out_n


def test_purity_partial_reductions_merge():
    """batch._merge_parts: the per-CTA partial rows of pfb_purity_metrics_pairs merge into the whole-series numbers."""
    from ska_pst_dsp_model_b200.batch import NPARTS, _merge_parts
    rng = np.random.default_rng(0)
    p = rng.random(1001)
    m = np.zeros((2, 3, NPARTS, 8))
    for k, ix in enumerate(np.array_split(np.arange(p.size), NPARTS)):
        v = p[ix]
        o = np.argsort(v)
        m[1, 2, k] = [v.sum(), v.max(), v.sum(), v[o[-1]], v[o[-2]], ix[o[-1]], np.sqrt(v).sum(), len(ix)]
    r = _merge_parts(m)[1, 2]
    o = np.argsort(p)
    assert np.allclose(r, [p.sum(), p.max(), p.sum(), p[o[-1]], p[o[-2]], o[-1], np.sqrt(p).sum(), p.size])


def test_filter_designs_against_independent_least_squares(tmp_path):
    """tools/design_filters.py: the firls designs (design_PFB_FIR_filter.m:24-48, design_PFB_FIR_filter_two_stage.m:42-78)
    agree with an independent weighted least-squares solve (Gauss-Legendre quadrature of the error integral + numpy
    lstsq on the cosine basis - no scipy.signal call), and the tool writes .mat / .txt / .npy with a SHA manifest that
    read_fir_filter_coeff loads back."""
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import design_filters as dfl
    from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter, read_fir_filter_coeff
    for nchan, nu, de, tpc in [(8, 8, 7, 10), (64, 4, 3, 8)]:
        h = design_pfb_fir_filter(nchan, nu, de, tpc)
        osf = nu / de
        ref = dfl.firls_reference(nchan * tpc + 1, [0, 1.0 / nchan, (2 * osf - 1) / nchan, 1.0], [1, 1, 0, 0], [1, 15])
        assert h.size == nchan * tpc + 1 and np.abs(h - ref).max() <= 1e-9 * np.abs(ref).max()
    # two-stage: the stage-1 design is the same least-squares problem at 32x the band edges (3137 taps); the long
    # filter is its FFT-domain interpolation by 32: same passband / stopband behaviour on the 32x finer grid
    from scipy.signal import firls, freqz
    osf = 8.0 / 7.0
    b1 = [0, 32.0 / 4096, 0.998 * 32 * (2 * osf - 1) / 4096, 1.0]
    h0 = dfl.firls_reference(3137, b1, [1, 1, 0, 0], [1, 15], grid_per_tap=3)
    assert np.abs(h0 - firls(3137, b1, [1, 1, 0, 0], weight=[1, 15])).max() < 1e-9 * np.abs(h0).max()
    hm, meta = dfl.design("two_stage", 4096, "8/7")
    assert hm.size == 100353 and abs(hm.sum() - 1.0) < 1e-3 and np.allclose(hm, hm[::-1], atol=1e-12)
    w = np.linspace(0, np.pi / 64, 257)                       # two coarse channels' worth of the long filter's band
    _, Hm = freqz(hm, 1, worN=w)
    _, H0 = freqz(h0, 1, worN=32 * w)
    assert np.abs(np.abs(Hm) - np.abs(H0)).max() < 2e-3
    rec = dfl.write(hm, meta, str(tmp_path), "Prototype_FIR.2_stage.8-7.4096.100352")
    man = json.load(open(tmp_path / "MANIFEST.json"))
    assert man["Prototype_FIR.2_stage.8-7.4096.100352"]["sha256_coefficients"] == rec["sha256_coefficients"]
    for ext in (".mat", ".txt", ".npy"):
        back = read_fir_filter_coeff(str(tmp_path / ("Prototype_FIR.2_stage.8-7.4096.100352" + ext)))
        assert np.array_equal(back, hm)


def test_maxflat_prototype_design():
    """generate_MaxFlt.m:32-71 restated (filters.generate_maxflat): the trigonometric resampler is exact on band-limited
    series in both directions, the Herrmann filter is symmetric, monotone, has unit gain at DC, a zero at Nyquist and its
    half-power point next to the requested frequency, and the ten passes drive the two-channel power sum to a constant."""
    from ska_pst_dsp_model_b200.filters import generate_maxflat, interpft, maxflat_sym
    k16, k64 = np.arange(16), np.arange(64)
    assert np.abs(interpft(np.cos(2 * np.pi * 3 * k16 / 16 + 0.3), 64) - np.cos(2 * np.pi * 3 * k64 / 64 + 0.3)).max() < 1e-12
    k25, k24 = np.arange(25), np.arange(24)
    assert np.abs(interpft(np.cos(2 * np.pi * 2 * k25 / 25 + 0.1), 24) - np.cos(2 * np.pi * 2 * k24 / 24 + 0.1)).max() < 1e-12
    h = maxflat_sym(24, 0.48)
    assert h.size == 25 and np.allclose(h, h[::-1], atol=1e-15) and abs(h.sum() - 1) < 1e-12
    assert abs(np.polyval(h, -1.0)) < 1e-12                                   # zero at z = -1
    H = np.abs(np.fft.rfft(h, 4096))
    assert np.all(np.diff(H) <= 1e-12)
    assert abs(np.argmin(np.abs(H - np.sqrt(0.5))) / 2048 - 0.48) < 0.02
    res = [generate_maxflat(32, 12, it)[1] for it in (0, 1, 3, 10)]
    assert res[0] > 1e-2 and all(a > b for a, b in zip(res, res[1:])) and res[-1] < 1e-10
    W, _ = generate_maxflat(32, 12)
    assert W.shape == (32 * 12,) and np.isfinite(W).all()
