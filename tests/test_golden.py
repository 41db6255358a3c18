"""Committed fixtures (tests/golden, made by make_golden.py from the oracle): the oracle must keep
reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os

import numpy as np
import pytest

from oracle import pfb_oracle as o

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def nerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _h(pst_taps):
    return pst_taps / pst_taps.sum()


def test_oracle_reproduces_golden(pst_taps):
    h = _h(pst_taps)
    g = load("analysis_8ch_8-7.npz")
    assert nerr(o.polyphase_analysis(g["x"].astype(np.complex128), h, 8, "8/7"), g["y"]) < 2e-7
    g = load("analysis_256ch_4-3.npz")
    assert nerr(o.polyphase_analysis_literal(g["x"].astype(np.complex128), h, 256, "4/3"), g["y"]) < 2e-7
    g = load("analysis_padded_16ch_8-7.npz")
    assert nerr(o.polyphase_analysis_padded_literal(g["x"].astype(np.complex128), g["taps"], 16, "8/7"), g["y"]) < 2e-7
    g = load("analysis_lowcbf.npz")
    assert nerr(o.polyphase_analysis_lowcbf(g["x"].astype(np.complex128), pst_taps, literal=True), g["y"]) < 2e-7
    g = load("inversion_8ch_8-7.npz")
    assert nerr(o.polyphase_synthesis_literal(g["x"].astype(np.complex128), True, 64, "8/7", g["taps"], 1, 8,
                                              o.window_tukey(64, 8)), g["y"]) < 2e-7
    g = load("inversion_64ch_4-3.npz")
    assert nerr(o.polyphase_synthesis_literal(g["x"].astype(np.complex128), True, 256, "4/3", g["taps"], 1, 48,
                                              o.window_tukey(256, 48)), g["y"]) < 2e-7


def _dev(x):
    import torch
    a = np.ascontiguousarray(np.transpose(x, (2, 1, 0)))
    if x.shape[1] == 1:
        a = a[:, 0, :]
    return torch.from_numpy(a).cuda()


def _host(t):
    a = t.cpu().numpy()
    if a.ndim == 2:
        a = a[:, None, :]
    return np.transpose(a, (2, 1, 0))


@pytest.mark.gpu
def test_cuda_matches_golden(pst_taps):
    from ska_pst_dsp_model_b200 import ops
    h = _h(pst_taps)
    g = load("analysis_8ch_8-7.npz")
    assert nerr(_host(ops.AnalysisPlan(8, "8/7", h)(_dev(g["x"]))), g["y"]) < 1e-5
    g = load("analysis_256ch_4-3.npz")
    assert nerr(_host(ops.AnalysisPlan(256, "4/3", h)(_dev(g["x"]))), g["y"]) < 1e-5
    g = load("analysis_padded_16ch_8-7.npz")
    assert nerr(_host(ops.AnalysisPlan(16, "8/7", g["taps"], "polyphase_analysis_padded")(_dev(g["x"]))), g["y"]) < 1e-5
    g = load("analysis_lowcbf.npz")
    assert nerr(_host(ops.AnalysisPlan(256, "4/3", pst_taps, "polyphase_analysis_lowcbf")(_dev(g["x"]))), g["y"]) < 1e-5
    g = load("inversion_8ch_8-7.npz")
    assert nerr(_host(ops.InversePlan(8, "8/7", 64, 8, "tukey", deripple_taps=g["taps"])(_dev(g["x"]))), g["y"]) < 1e-5
    g = load("inversion_64ch_4-3.npz")
    assert nerr(_host(ops.InversePlan(64, "4/3", 256, 48, "tukey", deripple_taps=g["taps"])(_dev(g["x"]))), g["y"]) < 1e-5
