#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the numpy oracle (NOT from the reference, which cannot
run here - see DESIGN.md §2).  The fixtures freeze the oracle against regressions and give the
GPU tests inputs/outputs that do not depend on the oracle code at test time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pfb_oracle as o  # noqa: E402

taps = np.loadtxt(os.path.join(HERE, "..", "..", "ska_pst_dsp_model_b200", "data", "pst_lowcbf_taps_int.txt"))
h = taps / taps.sum()
rng = np.random.default_rng(20261018)


def cn(shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


# 1. cfg1 shape (8 channels, OS 8/7, PST taps): impulse + noise, 40 blocks
x = cn((1, 1, 3072 + 7 * 40))
x[0, 0, 1700] += 50
np.savez_compressed(os.path.join(HERE, "analysis_8ch_8-7.npz"), x=x,
                    y=o.polyphase_analysis(x.astype(np.complex128), h, 8, "8/7").astype(np.complex64))
# 2. Low shape (256 channels, OS 4/3), dual pol, 6 blocks
x = cn((2, 1, 3072 + 192 * 6))
np.savez_compressed(os.path.join(HERE, "analysis_256ch_4-3.npz"), x=x,
                    y=o.polyphase_analysis(x.astype(np.complex128), h, 256, "4/3").astype(np.complex64))
# 3. padded variant, 16 channels 8/7, 401 taps
hp = (rng.standard_normal(401) / 20).astype(np.float64)
x = cn((2, 1, 14 * 60))
np.savez_compressed(os.path.join(HERE, "analysis_padded_16ch_8-7.npz"), x=x, taps=hp,
                    y=o.polyphase_analysis_padded(x.astype(np.complex128), hp, 16, "8/7").astype(np.complex64))
# 4. lowcbf, first call, 5 blocks
x = cn((1, 1, 1536 + 192 * 5))
np.savez_compressed(os.path.join(HERE, "analysis_lowcbf.npz"), x=x,
                    y=o.polyphase_analysis_lowcbf(x.astype(np.complex128), taps).astype(np.complex64))
# 5. inversion: 8 channels 8/7, Nf=64, ov=8, tukey, deripple with an 8-ch prototype
h8 = o.design_pfb_fir_filter(8, "8/7", 10)
x = cn((2, 8, 48 * 5 + 16))
np.savez_compressed(os.path.join(HERE, "inversion_8ch_8-7.npz"), x=x, taps=h8,
                    y=o.polyphase_synthesis(x.astype(np.complex128), True, 64, "8/7", h8, 1, 8,
                                            o.window_tukey(64, 8)).astype(np.complex64))
# 6. inversion: Low transform sizes (Nf=256, OS 4/3) on 64 channels, two overlap-save blocks
h64 = o.design_pfb_fir_filter(64, "4/3", 8)
x = cn((1, 64, 160 * 2 + 96))
np.savez_compressed(os.path.join(HERE, "inversion_64ch_4-3.npz"), x=x, taps=h64,
                    y=o.polyphase_synthesis(x.astype(np.complex128), True, 256, "4/3", h64, 1, 48,
                                            o.window_tukey(256, 48)).astype(np.complex64))
for f in sorted(os.listdir(HERE)):
    if f.endswith(".npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
