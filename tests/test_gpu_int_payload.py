"""Integer DADA payloads (NBIT 8 / 16: matlab/DADARead.m:41-53,66-70, write_dada_header.m:15-22) end to end:
device-side unpack / pack through the C ABI, and the streaming file operators on int8 / int16 files."""
import ctypes

import numpy as np
import pytest

from oracle import pfb_oracle as o

pytestmark = pytest.mark.gpu


def nerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("nbit", [8, 16])
@pytest.mark.parametrize("cdt", ["complex64", "complex128"])
def test_unpack_pack_bit_exact(nbit, cdt):
    import torch
    from ska_pst_dsp_model_b200 import lib as L
    rng = np.random.default_rng(nbit)
    idt = np.int8 if nbit == 8 else np.int16
    lim = 127 if nbit == 8 else 32767
    n = 100003                                                # odd length: exercises the scalar tail
    v = rng.integers(-lim - 1, lim + 1, 2 * n).astype(idt)
    vd = torch.from_numpy(v).cuda()
    tdt = getattr(torch, cdt)
    out = torch.empty(n, dtype=tdt, device="cuda")
    code = L.PFB_DTYPE_C64 if cdt == "complex64" else L.PFB_DTYPE_C128
    L.check(L.lib.pfb_unpack_payload(vd.data_ptr(), out.data_ptr(), 2 * n, nbit, code, None))
    got = out.cpu().numpy()
    ref = (v[0::2].astype(np.float64) + 1j * v[1::2].astype(np.float64)).astype(cdt)       # DADARead.m:66-70
    assert np.array_equal(got, ref)
    # pack: Matlab int16(scale * x): round half away from zero, saturate
    x = (rng.standard_normal(2 * n) * lim / 2).astype(np.float32 if cdt == "complex64" else np.float64)
    x[:8] = [0.5, -0.5, 1.5, -2.5, 1e9, -1e9, 2.4999, -0.4999]
    xd = torch.from_numpy(x).cuda()
    pk = torch.empty(2 * n, dtype=torch.int8 if nbit == 8 else torch.int16, device="cuda")
    L.check(L.lib.pfb_pack_payload(xd.data_ptr(), pk.data_ptr(), 2 * n, nbit, code, 1.0, None))
    want = np.clip(o.matlab_round(x.astype(np.float64)), -lim - 1, lim).astype(idt)
    assert np.array_equal(pk.cpu().numpy(), want)
    with pytest.raises(L.PfbError):
        L.check(L.lib.pfb_unpack_payload(vd.data_ptr(), out.data_ptr(), 2 * n, 12, code, None))


@pytest.mark.parametrize("nbit", [8, 16])
def test_streaming_operators_on_integer_files(tmp_path, pst_taps, nbit):
    """An NBIT 8 / 16 input file is channelized by the streaming operator (integers cross disk -> pinned memory ->
    PCIe, cast on the device) exactly as the whole-file operator that casts on the host, and both match the oracle
    fed the same integers (DADARead.m casts to double)."""
    from ska_pst_dsp_model_b200.file_streaming import stream_channelize, stream_synthesize
    from ska_pst_dsp_model_b200.format_handler import PSRFormatChannelizer, PSRFormatSynthesizer
    from ska_pst_dsp_model_b200.formats import DADAFile, format_header, DEFAULT_HEADER
    from ska_pst_dsp_model_b200.fft_windows import tukey_window
    rng = np.random.default_rng(7 + nbit)
    idt = np.int8 if nbit == 8 else np.int16
    lim = 100 if nbit == 8 else 20000
    n, npol = 3072 + 192 * 1200 + 5, 2
    raw = rng.integers(-lim, lim, (n, 1, npol, 2)).astype(idt)
    hdr = dict(DEFAULT_HEADER)
    hdr.update(NBIT=str(nbit), NDIM="2", NCHAN="1", NPOL=str(npol))
    path = tmp_path / f"in{nbit}.dump"
    with open(path, "wb") as f:
        f.write(format_header(hdr))
        raw.tofile(f)
    h = pst_taps / pst_taps.sum()
    chz = PSRFormatChannelizer("4/3", 256, h)
    whole = chz(DADAFile(str(path)).load_data(), output_dir=str(tmp_path), output_file_name="whole.dump")
    piece = stream_channelize(chz, str(path), output_dir=str(tmp_path), output_file_name="stream.dump", blocks_per_piece=304)
    a, b = open(whole.file_path, "rb").read(), open(piece.file_path, "rb").read()
    assert len(a) == len(b) and a == b
    assert piece["NBIT"] == "32" and piece.data.shape == (1200, 256, npol)          # lazily mapped payload
    x = (raw[..., 0].astype(np.float64) + 1j * raw[..., 1].astype(np.float64))      # (n, 1, npol)
    ref = o.polyphase_analysis(np.ascontiguousarray(np.transpose(x, (2, 1, 0))), h, 256, "4/3")
    assert nerr(np.transpose(np.asarray(piece.data), (2, 1, 0)), ref) < 1e-5
    # inversion of an integer channelized file (FilterBank.m:106-113 output rounding, stored as NBIT 16)
    ch = np.asarray(piece.data)
    scale = 30000.0 / np.abs(np.concatenate([ch.real.ravel(), ch.imag.ravel()])).max()
    chi = np.stack([o.matlab_round(ch.real * scale), o.matlab_round(ch.imag * scale)], axis=-1).astype(np.int16)
    hdr2 = dict(piece.header)
    hdr2.update(NBIT="16")
    p2 = tmp_path / "ch16.dump"
    with open(p2, "wb") as f:
        f.write(format_header(hdr2))
        chi.tofile(f)
    syn = PSRFormatSynthesizer(48, tukey_window(256, 48), 256, apply_deripple=True)
    s1 = syn(DADAFile(str(p2)).load_data(), output_dir=str(tmp_path), output_file_name="whole.syn.dump")
    s2 = stream_synthesize(syn, str(p2), output_dir=str(tmp_path), output_file_name="stream.syn.dump", blocks_per_piece=2)
    a, b = open(s1.file_path, "rb").read(), open(s2.file_path, "rb").read()
    assert len(a) == len(b) and a == b


def test_host_round_trip_with_integer_buffers(pst_taps):
    """streaming.HostRoundTrip fed int16 host samples and returning int16: identical to the float path on the same
    values (the device cast is exact) and to Matlab-rounding the float result on the host."""
    import torch
    from ska_pst_dsp_model_b200.streaming import HostRoundTrip
    h = pst_taps / pst_taps.sum()
    rng = np.random.default_rng(3)
    n, npol = 3072 + 192 * 3000, 2
    xi = rng.integers(-3000, 3000, (n, npol, 2)).astype(np.int16)
    xf = (xi[..., 0].astype(np.float32) + 1j * xi[..., 1].astype(np.float32)).astype(np.complex64)
    f = HostRoundTrip(n, 256, "4/3", h, 256, 48, "tukey", True, npol, nseg=4)
    yf = torch.empty((f.n_out, npol), dtype=torch.complex64).pin_memory()
    f.run(torch.from_numpy(xf).pin_memory(), yf)
    torch.cuda.synchronize()
    g = HostRoundTrip(n, 256, "4/3", h, 256, 48, "tukey", True, npol, nseg=4, in_nbit=16, out_nbit=16, out_scale=4.0)
    yi = torch.empty((g.n_out, npol, 2), dtype=torch.int16).pin_memory()
    g.run(torch.from_numpy(xi).pin_memory(), yi)
    g.run(torch.from_numpy(xi).pin_memory(), yi)          # second call replays the captured graph
    torch.cuda.synchronize()
    yr = yf.numpy()
    want = np.stack([o.matlab_round(4.0 * yr.real.astype(np.float64)), o.matlab_round(4.0 * yr.imag.astype(np.float64))], -1)
    # fp32 product 4*y is exact (power of two), so the device rounding sees the same numbers
    assert np.array_equal(yi.numpy(), np.clip(want, -32768, 32767).astype(np.int16))
