"""Drop-in replacements for the mcc-built executables the reference's ``backend="matlab"`` path runs
(python/data_gen/channelize.py:66-69, synthesize.py:67-71, generate_test_vector.py:171-174).  Same
positional argv (all strings), same exit-status contract (non-zero -> the caller raises
RuntimeError("Exited with non zero status"), python/data_gen/util.py:88-89), data in and out as DADA files.

    build/channelize <in.dump> <channels> <nu/de> <fir file> <out_name> <out_dir> <verbose> <use_padded>
    build/synthesize <in.dump> <Nf> <out_name> <out_dir> <verbose> <sample_offset> <deripple> <overlap> <window>
    build/generate_test_vector <handler> <n_bins> <p1,p2,..> <dtype> <n_pol> <header.json> <out_name> <out_dir> <verbose>

``ska_pst_dsp_model_b200/bin/*`` are the executables; link or copy them into the reference's ``build/``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np


def _arg(argv, i, default):
    return argv[i] if len(argv) > i else default


def channelize_main(argv):
    """matlab/channelize.m:32-111"""
    from .format_handler import PSRFormatChannelizer
    from .formats import DADAFile
    if len(argv) < 5:
        raise SystemExit("usage: channelize <in.dump> <channels> <nu/de> <fir file> <out_name> [out_dir] [verbose] [use_padded]")
    in_path, channels, os_str, fir_path, out_name = argv[:5]
    out_dir, verbose, use_padded = _arg(argv, 5, "./"), int(_arg(argv, 6, "0")), int(_arg(argv, 7, "0"))
    if verbose:
        print(f"channelize: input_file_path={in_path}\nchannelize: channels={channels} os_factor={os_str} "
              f"use_padded={use_padded}")
    ch = PSRFormatChannelizer(os_factor=os_str, nchan=int(channels), fir_filter_coeff=fir_path,
                              analysis_function="polyphase_analysis_padded" if use_padded else "polyphase_analysis")
    out = ch(DADAFile(in_path).load_data(), output_dir=out_dir, output_file_name=out_name)
    if verbose:
        print(f"channelize: saving channelized data complete: {out.file_path} {out.data.shape}")
    return 0


def synthesize_main(argv):
    """matlab/synthesize.m:26-111 (with the semantics of test_data_pipeline.m:130-132: the input spans the
    Nyquist zone; the stale positional call of synthesize.m:94-96 is not reproduced)."""
    from .fft_windows import make_window
    from .format_handler import PSRFormatSynthesizer
    from .formats import DADAFile
    if len(argv) < 3:
        raise SystemExit("usage: synthesize <in.dump> <Nf> <out_name> [out_dir] [verbose] [sample_offset] [deripple] [overlap] [window]")
    in_path, nfft, out_name = argv[0], int(argv[1]), argv[2]
    out_dir, verbose = _arg(argv, 3, "./"), int(_arg(argv, 4, "0"))
    sample_offset, deripple = int(_arg(argv, 5, "1")), int(_arg(argv, 6, "1"))
    overlap, window = int(_arg(argv, 7, "0")), _arg(argv, 8, "tukey")
    f = DADAFile(in_path).load_data()
    if sample_offset > 1:
        f.data = f.data[sample_offset - 1:]
    syn = PSRFormatSynthesizer(input_overlap=overlap, fft_window=make_window(window, nfft, overlap),
                               input_fft_length=nfft, apply_deripple=bool(deripple))
    out = syn(f, output_dir=out_dir, output_file_name=out_name)
    if verbose:
        print(f"synthesize: saving synthesized data complete: {out.file_path} {out.data.shape}")
    return 0


def generate_test_vector_main(argv):
    """matlab/generate_test_vector.m:32-111: params{1} is a fraction of n_bins; time_domain_impulse offsets are
    1-based (time_domain_impulse.m:18-25)."""
    from .formats import DADAFile
    if len(argv) < 7:
        raise SystemExit("usage: generate_test_vector <handler> <n_bins> <params> <dtype> <n_pol> <header.json> <out_name> [out_dir] [verbose]")
    handler, n_bins, params, dtype, n_pol, header_template, out_name = argv[:7]
    out_dir = _arg(argv, 7, "./")
    n_bins, n_pol = int(n_bins), int(n_pol)
    p = [float(v) for v in params.split(",")]
    p[0] = n_bins * p[0]
    cdt = np.complex64 if dtype == "single" else np.complex128
    t = np.arange(n_bins)
    if handler == "complex_sinusoid":
        phase = p[1] if len(p) > 1 else 0.0
        bin_offset = p[2] if len(p) > 2 else 0.0
        sig = np.exp(1j * (2 * np.pi * (p[0] + bin_offset) / n_bins * t + phase)).astype(cdt)
    elif handler == "time_domain_impulse":
        o, w = int(round(p[0])), int(p[1]) if len(p) > 1 else 1
        sig = np.zeros(n_bins, dtype=cdt)
        sig[o - 1: o - 1 + w] = 1.0
    else:
        raise SystemExit(f"generate_test_vector: unknown handler {handler}")
    f = DADAFile(os.path.join(out_dir, out_name))
    if os.path.exists(header_template):
        with open(header_template) as fh:
            f.header = json.load(fh)
    f.data = np.repeat(sig[:, None, None], n_pol, axis=2)
    f.dump_data()
    return 0


def main(prog=None, argv=None):
    argv = sys.argv[1:] if argv is None else argv
    prog = prog or os.path.basename(sys.argv[0])
    table = {"channelize": channelize_main, "synthesize": synthesize_main,
             "generate_test_vector": generate_test_vector_main}
    if prog not in table:
        if argv and argv[0] in table:
            prog, argv = argv[0], argv[1:]
        else:
            raise SystemExit(f"usage: {prog} {{channelize|synthesize|generate_test_vector}} args...")
    try:
        return table[prog](argv)
    except SystemExit:
        raise
    except Exception as e:      # noqa: BLE001 - the caller only sees the exit status and the log
        print(f"{prog}: error: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
