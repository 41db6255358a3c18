#!/usr/bin/env python
"""Prototype-filter designer: re-creates the coefficient files the reference's configs name but
git-ignores (config/Prototype_FIR.*.mat), as .mat (field `h`, plus n_chan / Fp / Fs like design_PFB_FIR_filter.m:43),
.txt and .npy, and records their SHA-256 in a manifest.

  single stage : matlab/design_PFB_FIR_filter.m:24-48          firls, order N*taps_per_chan, bands [0, 1/N, (2 OS-1)/N, 1],
                                                               weights [1, 15]
  two stage    : matlab/design_PFB_FIR_filter_two_stage.m:42-78 firls at 32x the band edges (0.998 Fs), then FFT-domain
                                                               zero stuffing to the full length (100 353 taps for Mid)
  maxflat      : matlab/generate_MaxFlt.m:32-71                 symmetric maximally flat two-channel filter, ten passes towards
                                                               constant total power, interpolated to channels x taps

    python tools/design_filters.py --out ska_pst_dsp_model_b200/data/designed            # all designs of test.config.json
    python tools/design_filters.py --design single_stage --channels 256 --os 4/3 --taps-per-chan 12 --out /tmp/f

`firls_reference` is an INDEPENDENT weighted least-squares solve (dense-grid quadrature + numpy lstsq on the cosine basis
of a type-I linear-phase filter) used by tests/test_host_logic.py to check the scipy.signal.firls call the designs rest on."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def firls_reference(numtaps, bands, desired, weight, grid_per_tap=24):
    """Weighted least squares on a dense frequency grid: minimise sum_b W_b * integral |A(f) - D(f)|^2 df over the bands,
    A(f) = a_0 + 2 sum_k a_k cos(pi k f) (f in units of Nyquist, as scipy.signal.firls), D linear inside each band."""
    assert numtaps % 2 == 1
    M = (numtaps - 1) // 2
    bands = np.asarray(bands, dtype=np.float64).reshape(-1, 2)
    desired = np.asarray(desired, dtype=np.float64).reshape(-1, 2)
    rows, rhs = [], []
    k = np.arange(M + 1)
    for (f0, f1), (d0, d1), w in zip(bands, desired, weight):
        n = max(64, int(np.ceil((f1 - f0) * numtaps * grid_per_tap)))
        # Gauss-Legendre nodes: the integrand is a low-degree trigonometric polynomial per sub-interval
        x, wq = np.polynomial.legendre.leggauss(8)
        edges = np.linspace(f0, f1, n + 1)
        mid, half = 0.5 * (edges[1:] + edges[:-1]), 0.5 * (edges[1:] - edges[:-1])
        f = (mid[:, None] + half[:, None] * x[None, :]).ravel()
        q = (half[:, None] * wq[None, :]).ravel() * w
        basis = np.cos(np.pi * f[:, None] * k[None, :])
        basis[:, 1:] *= 2.0
        d = d0 + (d1 - d0) * (f - f0) / (f1 - f0)
        sq = np.sqrt(q)
        rows.append(basis * sq[:, None])
        rhs.append(d * sq)
    a, *_ = np.linalg.lstsq(np.vstack(rows), np.concatenate(rhs), rcond=None)
    return np.concatenate([a[:0:-1], a])


def design(kind, channels, os_factor, taps_per_chan=None):
    from ska_pst_dsp_model_b200.filters import design_pfb_fir_filter, design_pfb_fir_filter_two_stage
    nu, de = (int(v) for v in str(os_factor).split("/"))
    osf = nu / de if nu != de else 1.1
    meta = {"n_chan": channels, "Fp": 1.0 / channels, "Fs": (2 * osf - 1) / channels, "os_factor": f"{nu}/{de}", "design": kind}
    if kind == "single_stage":
        h = design_pfb_fir_filter(channels, nu, de, taps_per_chan)
    elif kind == "two_stage":
        h = design_pfb_fir_filter_two_stage(channels, nu, de)
    elif kind == "maxflat":
        # matlab/generate_MaxFlt.m:32-71 (nbuff = channels, nTap = taps per channel): no band edges, the oversampling
        # factor plays no part in the design
        from ska_pst_dsp_model_b200.filters import generate_maxflat
        h, resid = generate_maxflat(channels, taps_per_chan)
        meta = {"n_chan": channels, "Fp": 0.0, "Fs": 0.0, "os_factor": f"{nu}/{de}", "design": kind,
                "power_complementarity_residual": resid}
    else:
        raise ValueError(f"unknown design {kind!r}")
    return np.asarray(h, dtype=np.float64), meta


def write(h, meta, out_dir, stem):
    from scipy.io import savemat
    os.makedirs(out_dir, exist_ok=True)
    paths = {}
    savemat(os.path.join(out_dir, stem + ".mat"), {"h": h.reshape(1, -1), "n_chan": float(meta["n_chan"]),
                                                    "Fp": meta["Fp"], "Fs": meta["Fs"]})
    np.savetxt(os.path.join(out_dir, stem + ".txt"), h, fmt="%.17e")
    np.save(os.path.join(out_dir, stem + ".npy"), h)
    for ext in (".mat", ".txt", ".npy"):
        with open(os.path.join(out_dir, stem + ext), "rb") as f:
            paths[stem + ext] = hashlib.sha256(f.read()).hexdigest()
    rec = dict(meta, ntaps=int(h.size), sum=float(h.sum()), sha256=paths,
               sha256_coefficients=hashlib.sha256(np.ascontiguousarray(h).tobytes()).hexdigest())
    man = os.path.join(out_dir, "MANIFEST.json")
    allrec = json.load(open(man)) if os.path.exists(man) else {}
    allrec[stem] = rec
    with open(man, "w") as f:
        json.dump(allrec, f, indent=1, sort_keys=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--design", choices=["single_stage", "two_stage", "maxflat"])
    ap.add_argument("--channels", type=int)
    ap.add_argument("--os", dest="os_factor")
    ap.add_argument("--taps-per-chan", type=int, default=12)
    args = ap.parse_args()
    jobs = []
    if args.design:
        jobs.append((args.design, args.channels, args.os_factor, args.taps_per_chan, None))
    else:       # every design named by the shipped configuration
        cfg = json.load(open(os.path.join(ROOT, "ska_pst_dsp_model_b200", "config", "test.config.json")))
        for name, c in cfg.items():
            if isinstance(c, dict) and "fir_filter_design" in c:
                tpc = c["fir_filter_taps"] // c["channels"]
                jobs.append((c["fir_filter_design"], c["channels"], c["os_factor"], tpc,
                             os.path.splitext(c["fir_filter_coeff_file_path"])[0]))
    for kind, ch, osf, tpc, stem in jobs:
        h, meta = design(kind, ch, osf, tpc)
        stem = stem or f"Prototype_FIR.{kind}.{osf.replace('/', '-')}.{ch}.{h.size - 1}"
        rec = write(h, meta, args.out, stem)
        print(stem, rec["ntaps"], "taps, sum", rec["sum"], rec["sha256_coefficients"][:16])


if __name__ == "__main__":
    main()
