"""BASELINE configs[4]: the batched purity sweep (impulse, tone, comb) on the device.  Every vector of a chunk goes
through ONE analysis launch over the concatenated chunk and ONE inversion launch pair; the results must equal the
oracle's per-vector generate -> polyphase_analysis -> polyphase_synthesis pipeline (purity.py:165-207)."""
import ctypes

import numpy as np
import pytest

from oracle import pfb_oracle as o

pytestmark = pytest.mark.gpu
TOL = 1e-5


def nerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def comb_numpy(n, shift_bins, phase, extra):
    """matlab/FrequencyComb.m:22-46 + sgcht.m:394-432: 32 tones, amplitudes linspace(1, sqrt 2, 32)."""
    t = np.arange(n, dtype=np.float64)
    amps = np.linspace(1.0, np.sqrt(2.0), 32)
    fmin = -0.5 + 1.0 / 128 + extra + shift_bins / n
    x = np.zeros(n, dtype=np.complex128)
    for k in range(32):
        x += amps[k] * np.exp(1j * (2 * np.pi * np.mod((fmin + k / 32.0) * t, 1.0) + phase))
    return x


def oracle_round_trip(x, h, taper):
    ch = o.polyphase_analysis(x[None, None, :].astype(np.complex128), h, 256, "4/3")
    return o.polyphase_synthesis(ch, True, 256, "4/3", h, 1, 48, taper)[0, 0]


@pytest.mark.parametrize("use_graph,fused", [(False, False), (True, True), (False, True)])
def test_batched_chunk_equals_per_vector_oracle(pst_taps, use_graph, fused):
    from ska_pst_dsp_model_b200.batch import BatchPurity
    h = pst_taps / pst_taps.sum()
    taper = o.window_tukey(256, 48)
    bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=6, use_graph=use_graph, fused=fused)
    assert bp.concat and bp.vec_rows == 768 and bp.nb_vec == 752 and bp.n_out == 4 * 30720 and bp.fused == fused
    n = bp.n_samples
    # --- impulses: 5 vectors (odd count: the spare slot repeats the last one), offsets incl. both ends
    offs = np.array([0, 12000, 40001, 70000, 140000])
    rep = bp.temporal(offs)
    assert [r["arg"] for r in rep] == [float(v) for v in offs]
    y = bp.y.cpu().numpy()
    for i, off in enumerate(offs):
        x = np.zeros(n, dtype=np.complex128)
        x[off] = 1.0
        ref = oracle_round_trip(x, h, taper)
        got = y[i >> 1, :, i & 1]
        assert got.shape == ref.shape
        if np.abs(ref).max() > 1e-3:            # (an impulse in the discarded head / tail leaves only leakage)
            assert nerr(got, ref) < TOL, (i, off)
        else:
            assert np.abs(got - ref).max() < TOL
        # the report's time-domain numbers are the reductions of exactly these series
        a = x[bp.total_sample_shift: bp.total_sample_shift + bp.n_cmp]
        d = np.abs(ref[: bp.n_cmp] - a)
        assert abs(rep[i]["total_diff"] - d.sum()) <= 2e-3 * d.sum() + 1e-4
    # --- tones
    freqs = np.array([3, 300, 49152 * 3 // 2, 147453])      # multiples of blocks = 3: whole cycles in fft_size (purity.py:96-99)
    rep_f = bp.spectral(freqs)
    y = bp.y.cpu().numpy()
    for i, f in enumerate(freqs):
        t = np.arange(n)
        x = np.exp(1j * (2 * np.pi * np.mod(f * t / n, 1.0) + np.pi / 4))
        ref = oracle_round_trip(x, h, taper)
        assert nerr(y[i >> 1, :, i & 1], ref) < TOL, (i, f)
        # spectral metrics against numpy's FFT of the oracle series (purity.py:236-263, verify/util.py:15-43)
        p = np.abs(np.fft.fft(ref[: bp.fft_size]) / bp.fft_size) ** 2
        p[np.argmax(p)] = 0.0
        assert abs(rep_f[i]["max_spurious_power"] - 10 * np.log10(p.max() + 1e-13)) < 0.05
        assert abs(rep_f[i]["total_spurious_power"] - 10 * np.log10(p.sum() + 1e-13)) < 0.05
        assert rep_f[i]["max_spurious_power"] < -60
    # --- combs
    shifts = np.array([0.0, 1000.25, 4607.5])
    rep_c = bp.comb(shifts)
    xg = bp.x.cpu().numpy()
    y = bp.y.cpu().numpy()
    for i, s in enumerate(shifts):
        x = comb_numpy(n, s, 0.0, 1.0 / 1024)
        assert nerr(xg[i >> 1, :, i & 1], x) < 2e-6                 # device generator == FrequencyComb definition
        ref = oracle_round_trip(x, h, taper)
        assert nerr(y[i >> 1, :, i & 1], ref) < TOL
        assert set(rep_c[i]) == {"mean_diff", "total_diff", "max_diff_power", "mean_signal_power", "arg"}


def test_own_spectrum_matches_numpy_fft():
    """pfb_spectrum_vectors (four-step 24 x 4096) against numpy.fft on random dual-pol series; bins come back
    transposed: out[k1][k2] = X[k1 + 24 k2]."""
    import torch
    from ska_pst_dsp_model_b200 import lib as L
    rng = np.random.default_rng(11)
    npairs, nfft, pitch = 3, 98304, 98304 + 777
    x = (rng.standard_normal((npairs, pitch, 2)) + 1j * rng.standard_normal((npairs, pitch, 2))).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    s1 = torch.empty((npairs, nfft, 2), dtype=torch.complex64, device="cuda")
    out = torch.empty_like(s1)
    L.check(L.lib.pfb_spectrum_vectors(xd.data_ptr(), pitch, s1.data_ptr(), out.data_ptr(), npairs, nfft, None))
    got = out.cpu().numpy().reshape(npairs, 24, 4096, 2)
    ref = np.fft.fft(x[:, :nfft, :].astype(np.complex128), axis=1)           # [pair][k][pol]
    ref_t = ref.reshape(npairs, 4096, 24, 2).transpose(0, 2, 1, 3)          # k = k1 + 24 k2 -> [k1][k2]
    assert nerr(got, ref_t) < 2e-6
    with pytest.raises(L.PfbError):
        L.check(L.lib.pfb_spectrum_vectors(xd.data_ptr(), pitch, s1.data_ptr(), out.data_ptr(), npairs, 65536, None))


def test_matlab_offset_grid(pst_taps):
    """current_performance.m:60-72 time_domain_offsets: the set contains the keep-length multiples from the filter
    offset and their +/- output-overlap neighbours."""
    from ska_pst_dsp_model_b200.batch import BatchPurity
    h = pst_taps / pst_taps.sum()
    bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=2, use_graph=False)
    p = bp.matlab_time_offsets(20)
    assert np.all(np.diff(p) >= 0) and p.min() >= 0 and p.max() < bp.n_samples
    fo, jump, ov = 1536, 49152 - 2 * 9216, 9216
    for v in (fo, fo + jump, fo + jump - ov, fo + ov, fo + 49152, 1):
        assert v in p
    rep = bp.temporal(p[:6])
    assert len(rep) == 6 and all(r["max_spurious_power"] < -60 for r in rep if 9216 + 1536 < r["arg"] < 100000)


def test_fused_round_trip_matches_oracle_and_two_plan_form(pst_taps):
    """ops.round_trip (pfb_roundtrip_execute): the analysis kernel writes the inversion's channel-IDFT intermediate itself
    (N x the rotated arm sums) - with and without the channelized output - and the result equals the oracle and the
    two-plan form to fp32 rounding."""
    import torch
    from ska_pst_dsp_model_b200 import ops
    from ska_pst_dsp_model_b200.lib import PfbError
    h = pst_taps / pst_taps.sum()
    rng = np.random.default_rng(17)
    n = 3072 + 192 * (16 * 70 + 5)
    x = (rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    ana = ops.AnalysisPlan(256, "4/3", h)
    inv = ops.InversePlan(256, "4/3", 256, 48, "tukey", deripple_taps=h)
    ch_ref = ana(xd)
    y_two = inv(ch_ref)
    ch, y1 = ops.round_trip(ana, inv, xd, want_chan=True)
    _, y2 = ops.round_trip(ana, inv, xd)
    assert torch.equal(ch, ch_ref)                                  # the channelized output is the same kernel code
    assert torch.equal(y1, y2)
    xo = np.ascontiguousarray(x.T[:, None, :]).astype(np.complex128)
    y_ref = o.polyphase_synthesis(o.polyphase_analysis(xo, h, 256, "4/3"), True, 256, "4/3", h, 1, 48, o.window_tukey(256, 48))
    got = y1.cpu().numpy().T[:, None, :]
    assert got.shape == y_ref.shape and nerr(got, y_ref) < TOL
    assert nerr(y_two.cpu().numpy(), y1.cpu().numpy()) < 2e-6
    with pytest.raises(PfbError):                                    # other shapes: not fused
        ops.round_trip(ops.AnalysisPlan(8, "8/7", h), ops.InversePlan(8, "8/7", 256, 48, "tukey"), xd[:, :1].contiguous())


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_pair_generator_equals_one_vector_generator(kind):
    """pfb_generate_test_vectors: the pair kernel (group == 2: both vectors of a [n][2] series from one CTA column,
    16-byte stores) writes exactly what the one-vector kernel writes (group == 1, one vector after the other)."""
    import torch
    from ska_pst_dsp_model_b200 import lib as _l
    n, nvec = 40000, 6
    rng = np.random.default_rng(kind)
    if kind == 0:
        prm = np.stack([rng.integers(0, n, nvec), np.full(nvec, 3), np.zeros(nvec)], axis=1).astype(np.float64)
    else:
        prm = np.stack([rng.uniform(0, n / 2, nvec), rng.uniform(0, 1, nvec), np.full(nvec, 1.0 / 1024 if kind == 2 else 0.25)],
                       axis=1)
    dprm = torch.from_numpy(prm).cuda()
    pairs = torch.zeros((nvec // 2, n, 2), dtype=torch.complex64, device="cuda")
    singles = torch.zeros((nvec, n), dtype=torch.complex64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _l.check(_l.lib.pfb_generate_test_vectors(kind, pairs.data_ptr(), n, nvec, 2, n * 2, _l.PFB_DTYPE_C64, dprm.data_ptr(), st))
    _l.check(_l.lib.pfb_generate_test_vectors(kind, singles.data_ptr(), n, nvec, 1, n, _l.PFB_DTYPE_C64, dprm.data_ptr(), st))
    torch.cuda.synchronize()
    got = pairs.permute(0, 2, 1).reshape(nvec, n)
    assert float((got - singles).abs().max()) <= 2e-7          # same factors; the span of the A/B tables differs
    assert float(singles.abs().max()) > 0


def test_async_sweep_equals_kind_by_kind_runs(pst_taps):
    """BatchPurity.sweep enqueues all kinds before it builds any report (pinned staging, reserve()); the reports equal the
    ones of the synchronous per-kind entry points."""
    import torch
    from ska_pst_dsp_model_b200.batch import BatchPurity
    h = pst_taps / pst_taps.sum()
    bp = BatchPurity("4/3", 256, 48, "tukey", True, 256, h, 3, vectors_per_batch=16)
    bp.reserve(40)
    rep = bp.sweep(40, 3)
    offsets = np.linspace(1, bp.n_samples, 40).astype(int)
    freqs = np.linspace(1, bp.block_size, 40).astype(int) * 3
    shifts = np.linspace(0, bp.n_samples / 32.0, 40)
    for got, ref in ((rep["test_time_domain_impulse"], bp.temporal(offsets)), (rep["test_complex_sinusoid"], bp.spectral(freqs)),
                     (rep["test_frequency_comb"], bp.comb(shifts))):
        assert len(got) == len(ref) == 40
        for a, b in zip(got, ref):
            assert a.keys() == b.keys()
            for k in a:
                assert a[k] == b[k] or abs(a[k] - b[k]) <= 1e-12 * max(1.0, abs(b[k]))
    torch.cuda.synchronize()
