"""GPU tests through the reference's Python call surface (files in, files out)."""
import os

import numpy as np
import pytest

from oracle import pfb_oracle as o

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_channelize_synthesize_files_low(tmp_path, pst_taps):
    from ska_pst_dsp_model_b200 import data_gen
    from ska_pst_dsp_model_b200.formats import DADAFile
    data_gen.set_config("low_pst")
    d = str(tmp_path)
    n = 256 * 192 * 3
    gen = data_gen.generate_test_vector(backend="python", n_bins=n, n_pol=2)
    src = gen([30000], [1], domain_name="time", output_dir=d)
    ch = data_gen.channelize(src.file_path, output_dir=d, backend="b200")
    assert os.path.basename(ch.file_path) == "channelize.256.4-3.dump"            # channelize.py:58-62
    assert ch.data.shape == ((n - 3072) // 192, 256, 2)
    assert ch["OS_FACTOR"] == "4/3" and ch["NCHAN"] == "256" and ch["PFB_DC_CHAN"] == "1" and ch["NCHAN_PFB_0"] == "256"
    assert ch["NTAP_0"] == "3072" and ch["NSTAGE"] == "1" and int(ch["HDR_SIZE"]) >= 65536
    assert abs(float(ch["TSAMP"]) - 0.025 * 256 * 3 / 4) < 1e-12
    h = pst_taps                      # raw integer taps, as the config file holds them
    xo = np.transpose(src.data, (2, 1, 0)).astype(np.complex128)
    ref = o.polyphase_analysis(xo, h, 256, "4/3")
    assert nerr(np.transpose(ch.data, (2, 1, 0)), ref) < 1e-5
    inv = data_gen.synthesize(ch.file_path, fft_window_str="tukey", deripple=True, backend="b200", output_dir=d)
    assert os.path.basename(inv.file_path) == "synthesize.256.dump"                 # synthesize.py:60-63
    # the synthesizer takes its deripple taps from COEFF_0 (%0.6E, add_fir_filter_to_header.m:14)
    coeff = np.array([float(v) for v in ch["COEFF_0"].split(",")])
    ref_inv = o.polyphase_synthesis(np.transpose(ch.data, (2, 1, 0)).astype(np.complex128), True, 256, "4/3", coeff, 1, 48,
                                    o.window_tukey(256, 48))
    got = np.transpose(DADAFile(inv.file_path).load_data().data, (2, 1, 0))
    assert got.shape == ref_inv.shape and nerr(got, ref_inv) < 1e-5
    assert abs(float(inv["TSAMP"]) - 0.025) < 1e-12


def test_channelize_cfg1_matches_oracle(tmp_path, pst_taps):
    """BASELINE configs[0]: 8 channels, OS 8/7, PST taps, 2^18 complex64 single-pol impulse."""
    from ska_pst_dsp_model_b200 import data_gen
    data_gen.set_config("test")
    d = str(tmp_path)
    n = 1 << 18
    src = data_gen.generate_test_vector([0.11], [1], n_bins=n, domain_name="time", backend="python", output_dir=d)
    ch = data_gen.channelize(backend="python")(src.file_path, channels=8, os_factor_str="8/7", output_dir=d)
    assert ch.data.shape == (37010, 8, 1)
    ref = o.polyphase_analysis(np.transpose(src.data, (2, 1, 0)).astype(np.complex128), pst_taps, 8, "8/7")
    assert nerr(np.transpose(ch.data, (2, 1, 0)), ref) < 1e-5


def test_backends_driver_with_oracle_backend(tmp_path, pst_taps):
    """verify.test_backends: the b200 channelizer against a second backend (here the oracle,
    registered from the test side) at the reference's atol = rtol = 1e-4."""
    from ska_pst_dsp_model_b200 import data_gen
    from ska_pst_dsp_model_b200.data_gen.channelize import register_backend
    from ska_pst_dsp_model_b200.filters import resolve_fir_filter
    from ska_pst_dsp_model_b200.formats import DADAFile
    from ska_pst_dsp_model_b200.verify.test_backends import TestBackends
    data_gen.set_config("low_pst")

    def oracle_backend(f, channels, os_str, fir_path, output_dir, output_file_name, use_padded):
        taps = resolve_fir_filter(fir_path)
        taps = taps / taps.sum()
        fn = o.polyphase_analysis_padded if use_padded else o.polyphase_analysis
        y = fn(np.transpose(f.data, (2, 1, 0)).astype(np.complex128), taps, channels, os_str)
        out = DADAFile(os.path.join(output_dir, output_file_name), header=f.header)
        out.data = np.transpose(y, (2, 1, 0)).astype(np.complex64)
        return out.dump_data()

    register_backend("oracle", oracle_backend)
    norm = tmp_path / "taps.npy"
    np.save(norm, pst_taps / pst_taps.sum())
    t = TestBackends("4/3", 256, 256, str(norm), blocks=2, backends=("b200", "oracle"), output_dir=str(tmp_path))
    res = t.test_channelizer_backends()
    assert res["oracle"]["mean"] == 1.0


def test_purity_driver_meets_requirement(tmp_path, pst_taps):
    """verify.purity on a handful of vectors: -60 dB temporal and spectral purity (TestImpulse.m / TestPureTone.m)."""
    from ska_pst_dsp_model_b200 import data_gen
    from ska_pst_dsp_model_b200.verify.purity import TestPurity
    data_gen.set_config("low_pst")
    norm = tmp_path / "taps.npy"
    np.save(norm, pst_taps / pst_taps.sum())
    t = TestPurity(n_test=4, os_factor="4/3", input_fft_length=256, input_overlap=48, fft_window="tukey", deripple=True,
                   channels=256, fir_filter_taps=3072, fir_filter_coeff_file_path=str(norm), blocks=3,
                   output_dir=str(tmp_path))
    # keep the offsets where an inverted impulse exists (the last offset lies beyond the inverted span)
    t.time_domain_args["offset"] = np.array([12000, 40000, 70000])
    t.temporal_purity()
    t.spectral_purity()
    worst = t.check(-60.0)
    assert set(worst) == {"test_time_domain_impulse", "test_complex_sinusoid"}
    path = t.finish()
    import json
    rep = json.load(open(path))
    assert set(rep["test_time_domain_impulse"][0]) == {"mean_diff", "total_diff", "max_spurious_power",
                                                        "total_spurious_power", "mean_spurious_power", "arg"}


def test_batch_purity_matches_file_driver(tmp_path, pst_taps):
    """The on-device batched sweep reproduces the file-based purity driver's metrics."""
    from ska_pst_dsp_model_b200 import data_gen
    from ska_pst_dsp_model_b200.batch import BatchPurity
    from ska_pst_dsp_model_b200.verify.purity import TestPurity
    data_gen.set_config("low_pst")
    h = pst_taps / pst_taps.sum()
    norm = tmp_path / "taps.npy"
    np.save(norm, h)
    t = TestPurity(n_test=4, os_factor="4/3", input_fft_length=256, input_overlap=48, fft_window="tukey", deripple=True,
                   channels=256, fir_filter_taps=3072, fir_filter_coeff_file_path=str(norm), blocks=3,
                   output_dir=str(tmp_path))
    offs, freqs = np.array([12000, 40000, 70000]), np.array([3, 300, 49152 * 3 // 2])
    t.time_domain_args["offset"], t.freq_domain_args["frequency"] = offs, freqs
    t.temporal_purity()
    t.spectral_purity()
    # the file driver reads the deripple taps back from the channelized file's COEFF_0 key (%0.6E, synthesize.m:83-85 /
    # add_fir_filter_to_header.m:14): give the device driver the same rounded taps so that both compute the same thing
    h_hdr = np.array([float("%0.6E" % c) for c in h])
    bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=2, deripple_taps=h_hdr)
    assert bp.total_sample_shift == t.total_sample_shift and bp.n_samples == t.n_samples
    got_t, got_f = bp.temporal(offs), bp.spectral(freqs)
    for key, got in (("test_time_domain_impulse", got_t), ("test_complex_sinusoid", got_f)):
        for r, g in zip(t.report[key], got):
            assert r["arg"] == g["arg"]
            assert abs(r["total_diff"] - g["total_diff"]) <= 2e-3 * abs(r["total_diff"]) + 1e-6
            for k in ("max_spurious_power", "total_spurious_power", "mean_spurious_power"):
                assert abs(r[k] - g[k]) < 0.05, (key, k, r[k], g[k])
            assert g["max_spurious_power"] < -60


def test_device_generators_match_reference_definitions():
    import ctypes
    import torch
    from ska_pst_dsp_model_b200 import data_gen
    from ska_pst_dsp_model_b200 import lib as L
    n = 5000
    x = torch.empty((n, 2), dtype=torch.complex64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib.pfb_generate_test_vector(0, x.data_ptr(), n, 2, 0, 1234.0, 3.0, 0.0, st))
    ref = data_gen.time_domain_impulse(n, [1234], [3])
    assert np.array_equal(x.cpu().numpy()[:, 0], ref) and np.array_equal(x.cpu().numpy()[:, 1], ref)
    L.check(L.lib.pfb_generate_test_vector(1, x.data_ptr(), n, 2, 0, 37.0, np.pi / 4, 0.1, st))
    ref = data_gen.complex_sinusoid(n, [37], [np.pi / 4], 0.1)
    assert np.abs(x.cpu().numpy()[:, 1] - ref).max() < 2e-7


def test_mid_config_round_trip_impulse(tmp_path):
    """BASELINE configs[2] (SKA-Mid): 4096 channels, OS 8/7, the re-created 100 353-tap two-stage prototype,
    polyphase_analysis_padded + inversion (Nf=512, overlap 128, tukey, deripple) through the call surface.
    TestImpulse criterion with (fir_offset_direction, kludge_offset) = (0, 0) of the 'mid' config:
    peak at offset - (output_overlap - 1), leakage < -60 dB; and parity with the oracle on the analysis."""
    from ska_pst_dsp_model_b200 import data_gen
    from ska_pst_dsp_model_b200.filters import resolve_fir_filter
    cfg = data_gen.set_config("mid")
    d = str(tmp_path)
    n = 3584 * (256 * 3 + 2 * 128 + 40)
    off = 3584 * 600 + 1234
    src = data_gen.generate_test_vector([off], [1], n_bins=n, domain_name="time", backend="python", output_dir=d)
    ch = data_gen.channelize(src.file_path, output_dir=d, backend="b200")     # analysis_function = padded (config)
    assert ch.data.shape == (n // 3584, 4096, 1) and ch["OS_FACTOR"] == "8/7" and ch["NTAP_0"] == "100353"
    taps = resolve_fir_filter(cfg["fir_filter_coeff_file_path"], cfg)
    ref = o.polyphase_analysis_padded(np.transpose(src.data[: 3584 * 700], (2, 1, 0)).astype(np.complex128), taps, 4096, "8/7")
    got = np.transpose(ch.data, (2, 1, 0))
    # away from the circular wrap of the delay compensation the prefix of the stream gives the same blocks
    assert nerr(got[:, :, 560:640], ref[:, :, 560:640]) < 1e-5
    inv = data_gen.synthesize(ch.file_path, fft_window_str="tukey", deripple=True, backend="b200", output_dir=d)
    y = inv.data[:, 0, 0]
    out_ov = 128 * 7 // 8 * 4096
    expect = off + 0 - (out_ov - 1 + 0)
    pk = int(np.argmax(np.abs(y)))
    assert abs(pk - expect) <= 1, (pk, expect)
    rest = np.abs(y).copy() / np.abs(y[pk])
    rest[pk - 1: pk + 2] = 0
    assert 20 * np.log10(rest.max()) < -60


def test_matlab_cli_replacements(tmp_path, pst_taps):
    """The reference's backend="matlab" path: subprocess of build/<exe> with positional argv (util.run_cmd)."""
    import subprocess
    import sys
    from ska_pst_dsp_model_b200.formats import DADAFile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = os.path.join(root, "ska_pst_dsp_model_b200", "bin")
    d = str(tmp_path)
    hdr = tmp_path / "hdr.json"
    hdr.write_text('{"HDR_SIZE": "4096", "TSAMP": "0.025", "NCHAN": "1", "NPOL": "2", "NBIT": "32", "NDIM": "2"}')
    taps = tmp_path / "taps.npy"
    np.save(taps, pst_taps / pst_taps.sum())
    n = 256 * 192 * 3
    run = lambda *a: subprocess.run([sys.executable] + [str(x) for x in a], check=False, capture_output=True, text=True)
    r = run(os.path.join(bindir, "generate_test_vector"), "time_domain_impulse", n, "0.25,1", "single", 2, hdr,
            "imp.dump", d, 1)
    assert r.returncode == 0, r.stderr
    src = DADAFile(os.path.join(d, "imp.dump")).load_data()
    assert src.data.shape == (n, 1, 2) and src.data[n // 4 - 1, 0, 0] == 1 and src.data.sum() == 2   # 1-based offset
    r = run(os.path.join(bindir, "channelize"), src.file_path, 256, "4/3", taps, "chan.dump", d, 1, 0)
    assert r.returncode == 0, r.stderr
    r = run(os.path.join(bindir, "synthesize"), os.path.join(d, "chan.dump"), 256, "inv.dump", d, 1, 1, 1, 48, "tukey")
    assert r.returncode == 0, r.stderr
    y = DADAFile(os.path.join(d, "inv.dump")).load_data().data[:, 0, 0]
    expect = (n // 4 - 1) - 1536 - 36 * 256
    assert int(np.argmax(np.abs(y))) == expect
    r = run(os.path.join(bindir, "channelize"), "/nonexistent.dump", 256, "4/3", taps, "x.dump", d)
    assert r.returncode != 0                                             # -> RuntimeError in the reference's run_cmd


def test_strided_generator_fills_one_polarisation_slot():
    """pfb_generate_test_vector_strided writes out[t * stride] only (two vectors share one [n][2] stream)."""
    import ctypes
    import torch
    from ska_pst_dsp_model_b200 import data_gen
    from ska_pst_dsp_model_b200 import lib as L
    n = 4096
    x = torch.full((n, 2), 7 + 7j, dtype=torch.complex64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib.pfb_generate_test_vector_strided(0, x.data_ptr() + 8, n, 2, 0, 100.0, 2.0, 0.0, st))
    got = x.cpu().numpy()
    assert np.all(got[:, 0] == 7 + 7j)
    assert np.array_equal(got[:, 1], data_gen.time_domain_impulse(n, [100], [2]))
    L.check(L.lib.pfb_generate_test_vector_strided(1, x.data_ptr(), n, 2, 0, 5.0, 0.3, 0.0, st))
    assert np.abs(x.cpu().numpy()[:, 0] - data_gen.complex_sinusoid(n, [5], [0.3], 0.0)).max() < 2e-7


def test_host_round_trip_graph_replay_matches_eager(pst_taps):
    """streaming.HostRoundTrip: the captured CUDA graph replays the same result as the eager segment loop and as the
    one-shot device-resident plans (segments are cut on inversion-block boundaries, so the match is exact)."""
    import torch
    from ska_pst_dsp_model_b200 import ops
    from ska_pst_dsp_model_b200.streaming import HostRoundTrip
    h = pst_taps / pst_taps.sum()
    rng = np.random.default_rng(11)
    n = 3072 + 192 * (160 * 24 + 96)
    x = torch.from_numpy((rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))).astype(np.complex64)).pin_memory()
    ref = ops.InversePlan(256, "4/3", 256, 48, "tukey", deripple_taps=h)(ops.AnalysisPlan(256, "4/3", h)(x.cuda())).cpu()
    for use_graph in (False, True):
        hrt = HostRoundTrip(n, 256, "4/3", h, 256, 48, "tukey", True, 2, nseg=4, use_graph=use_graph)
        y = torch.zeros((hrt.n_out, 2), dtype=torch.complex64).pin_memory()
        for _ in range(2):                       # second call replays the graph
            y.zero_()
            hrt.run(x, y)
            torch.cuda.synchronize()
            assert tuple(y.shape) == tuple(ref.shape)
            assert torch.equal(y, ref), use_graph


@pytest.mark.parametrize("variant", ["polyphase_analysis", "polyphase_analysis_lowcbf"])
def test_streaming_file_operators_match_whole_file(tmp_path, pst_taps, variant):
    """file_streaming.stream_channelize / stream_synthesize (pieces read into pinned memory, three streams, writer
    thread) produce byte-identical DADA files to the whole-file PSRFormatChannelizer / PSRFormatSynthesizer."""
    from ska_pst_dsp_model_b200.file_streaming import stream_channelize, stream_synthesize
    from ska_pst_dsp_model_b200.format_handler import PSRFormatChannelizer, PSRFormatSynthesizer
    from ska_pst_dsp_model_b200.formats import DADAFile
    from ska_pst_dsp_model_b200.fft_windows import tukey_window
    rng = np.random.default_rng(5)
    n = 3072 + 192 * 2500 + 77
    x = (rng.standard_normal((n, 1, 2)) + 1j * rng.standard_normal((n, 1, 2))).astype(np.complex64)
    src = DADAFile(str(tmp_path / "in.dump"), data=x)
    src.dump_data()
    h = pst_taps if variant.endswith("lowcbf") else pst_taps / pst_taps.sum()
    chz = PSRFormatChannelizer("4/3", 256, h, analysis_function=variant)
    whole = chz(DADAFile(src.file_path).load_data(), output_dir=str(tmp_path), output_file_name="whole.ch.dump")
    piecewise = stream_channelize(chz, src.file_path, output_dir=str(tmp_path), output_file_name="stream.ch.dump",
                                  blocks_per_piece=336)
    a, b = open(whole.file_path, "rb").read(), open(piecewise.file_path, "rb").read()
    assert len(a) == len(b) and a == b
    if variant.endswith("lowcbf"):
        return
    syn = PSRFormatSynthesizer(48, tukey_window(256, 48), 256, apply_deripple=True)
    whole_s = syn(DADAFile(whole.file_path).load_data(), output_dir=str(tmp_path), output_file_name="whole.syn.dump")
    piece_s = stream_synthesize(syn, whole.file_path, output_dir=str(tmp_path), output_file_name="stream.syn.dump",
                                blocks_per_piece=3)
    a, b = open(whole_s.file_path, "rb").read(), open(piece_s.file_path, "rb").read()
    assert len(a) == len(b) and a == b
