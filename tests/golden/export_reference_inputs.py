#!/usr/bin/env python
"""Writes tests/golden/reference_inputs.mat: the exact inputs, filter taps and parameters of every
tests/golden/*.npz fixture, in the layout the reference's Matlab functions take ((npol, nchan, ndat) arrays).

The reference cannot run in the build container (no Matlab / Octave).  Someone who has Matlab runs
tests/golden/make_reference_fixtures.m on this file; it calls the reference's own polyphase_analysis,
polyphase_analysis_padded, polyphase_analysis_lowcbf and polyphase_synthesis and writes
tests/golden/reference_outputs.mat, which tests/test_golden.py then uses to pin the oracle (and the CUDA path)
against the reference itself.

    python tests/golden/export_reference_inputs.py
"""
import os

import numpy as np
from scipy.io import savemat

HERE = os.path.dirname(os.path.abspath(__file__))
taps = np.loadtxt(os.path.join(HERE, "..", "..", "ska_pst_dsp_model_b200", "data", "pst_lowcbf_taps_int.txt"))
h = taps / taps.sum()


def g(name):
    return np.load(os.path.join(HERE, name))


cases = {}
f = g("analysis_8ch_8-7.npz")
cases["a8"] = dict(x=f["x"], filt=h, nchan=8, nu=8, de=7, func="polyphase_analysis")
f = g("analysis_256ch_4-3.npz")
cases["a256"] = dict(x=f["x"], filt=h, nchan=256, nu=4, de=3, func="polyphase_analysis")
f = g("analysis_padded_16ch_8-7.npz")
cases["ap16"] = dict(x=f["x"], filt=f["taps"], nchan=16, nu=8, de=7, func="polyphase_analysis_padded")
f = g("analysis_lowcbf.npz")
cases["alow"] = dict(x=f["x"], filt=taps, nchan=256, nu=4, de=3, func="polyphase_analysis_lowcbf")
f = g("inversion_8ch_8-7.npz")
cases["i8"] = dict(x=f["x"], filt=f["taps"], nchan=8, nu=8, de=7, nfft=64, overlap=8, func="polyphase_synthesis")
f = g("inversion_64ch_4-3.npz")
cases["i64"] = dict(x=f["x"], filt=f["taps"], nchan=64, nu=4, de=3, nfft=256, overlap=48, func="polyphase_synthesis")
out = {}
for k, c in cases.items():
    c = dict(c)
    c["x"] = np.asarray(c["x"], dtype=np.complex64)                  # (npol, nchan, ndat), single as default_config.m:23
    c["filt"] = np.asarray(c["filt"], dtype=np.float64).reshape(-1, 1)
    for key in ("nchan", "nu", "de", "nfft", "overlap"):
        if key in c:
            c[key] = float(c[key])
    out[k] = c
savemat(os.path.join(HERE, "reference_inputs.mat"), out, do_compression=True, oned_as="column")
print("wrote reference_inputs.mat", os.path.getsize(os.path.join(HERE, "reference_inputs.mat")), "bytes:", sorted(out))
