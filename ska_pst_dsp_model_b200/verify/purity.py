"""Temporal / spectral purity of the PFB inversion (python/verify/purity.py:31-301):
one generate -> channelize -> synthesize -> compare pass per test vector, spurious-power
metrics, JSON report with the reference's keys.

The un-vendored ``comparator`` package is replaced by the few operations purity.py uses
from it: time domain ``this`` / ``diff`` with |.| mean, sum, max products, and the
frequency domain as the FFT of the first ``fft_size`` samples (``comp.freq.domain =
[0, fft_size]``, purity.py:150).  Unlike the reference ``check()``
turns the -60 dB requirement into a hard assertion.
"""
from __future__ import annotations

import functools
import glob
import json
import logging
import os
import tempfile

import numpy as np

from .. import data_gen
from ..data_gen import util as dg_util
from ..rational import Rational
from . import util as test_util

module_logger = logging.getLogger(__name__)


class NumpyEncoder(json.JSONEncoder):
    def default(self, o):
        if isinstance(o, np.integer):
            return int(o)
        if isinstance(o, np.floating):
            return float(o)
        if isinstance(o, np.ndarray):
            return o.tolist()
        return super().default(o)


def _products(a):
    return {"mean": float(np.mean(np.abs(a))), "sum": float(np.sum(np.abs(a))), "max": float(np.amax(np.abs(a))),
            "total_spurious": float(test_util.total_spurious(a)), "mean_spurious": float(test_util.mean_spurious(a)),
            "max_spurious": float(test_util.max_spurious(a))}


class TestPurity:
    __test__ = False   # not a pytest class
    thresh = 1e-7
    time_domain_args = {"width": 1}
    freq_domain_args = {"phase": np.pi / 4, "bin_offset": 0.0}

    def __init__(self, n_test, os_factor, input_fft_length, input_overlap, fft_window, deripple, channels,
                 fir_filter_taps, fir_filter_coeff_file_path, blocks, backend=None, n_pol=1, output_dir=None,
                 products_dir=None, save_output=False, **_ignored):
        backend = backend or {"test_vectors": "python", "channelize": "b200", "synthesize": "b200"}
        self.input_fft_length, self.input_overlap = input_fft_length, input_overlap
        self.deripple, self.fft_window, self.save_output = deripple, fft_window, save_output
        self.output_dir = output_dir or tempfile.mkdtemp(prefix="pfb_purity_")
        self.products_dir = products_dir or self.output_dir
        os_factor = Rational.from_str(os_factor)
        block_size = os_factor.normalize(input_fft_length) * channels              # purity.py:82
        self.fft_size = 2 * block_size
        self.n_samples = block_size * blocks
        self.output_sample_shift = os_factor.normalize(input_overlap) * channels
        self.total_sample_shift = self.output_sample_shift + (fir_filter_taps - 1) // 2   # :86-88
        self.time_domain_args = dict(self.time_domain_args)
        self.freq_domain_args = dict(self.freq_domain_args)
        if n_test == 1:
            self.time_domain_args["offset"] = [10 + self.total_sample_shift]
            self.freq_domain_args["frequency"] = [1 * blocks]
        else:
            self.time_domain_args["offset"] = np.linspace(1, self.n_samples, n_test).astype(int)
            self.freq_domain_args["frequency"] = np.linspace(1, block_size, n_test).astype(int) * blocks
        self.block_size, self.os_factor = block_size, os_factor
        self.generator = data_gen.generate_test_vector(backend=backend["test_vectors"], n_bins=self.n_samples,
                                                       n_pol=n_pol)
        fir_path = fir_filter_coeff_file_path
        self.channelizer = functools.partial(data_gen.channelize(backend=backend["channelize"]), channels=channels,
                                             os_factor_str=str(os_factor), fir_filter_path=fir_path)
        self.pipeline = data_gen.pipeline(self.generator, self.channelizer, lambda a, **kwargs: a,
                                          output_dir=self.output_dir)
        synthesizer = functools.partial(data_gen.synthesize, deripple=deripple, backend=backend["synthesize"],
                                        fft_window_str=fft_window, input_fft_length=input_fft_length,
                                        input_overlap=input_overlap, output_dir=self.output_dir)
        self.synthesizer = lambda a: [synthesizer(a, output_file_name="synthesized." + os.path.basename(a))]
        self.report = {}
        self.files = []

    # purity.py:165-207
    def _test(self, *, test_vector_func, test_vector_args, test_method_name, report_func):
        method_report = []
        for arg in test_vector_args:
            dump_files = test_vector_func(arg)
            inverted_dump = self.synthesizer(dump_files[1].file_path)[0]
            input_dat, inverted_dat = self.chop(dump_files[0], inverted_dump)
            n = min(input_dat.size, inverted_dat.size)
            a, b = input_dat[:n], inverted_dat[:n]
            time_prod = {"this": [_products(a), _products(b)], "diff": _products(b - a)}
            fa = np.fft.fft((a / self.fft_size)[: self.fft_size])
            fb = np.fft.fft((b / self.fft_size)[: self.fft_size])
            freq_prod = {"this": [_products(fa), _products(fb)], "diff": _products(fb - fa)}
            sub_report = report_func(time_prod, freq_prod)
            sub_report["arg"] = arg
            method_report.append(sub_report)
            self.files.extend(dump_files)
            self.files.append(inverted_dump)
            if not self.save_output:
                self.dispose()
            self.report[test_method_name] = method_report

    def temporal_purity(self):
        func = dg_util.rpartial(functools.partial(self.pipeline, domain_name="time"), self.time_domain_args["width"])

        def report(time_prod, freq_prod):
            d, t = time_prod["diff"], time_prod["this"][1]
            return {"mean_diff": d["mean"], "total_diff": d["sum"], "max_spurious_power": t["max_spurious"],
                    "total_spurious_power": t["total_spurious"], "mean_spurious_power": t["mean_spurious"]}

        self._test(test_vector_func=func, test_vector_args=self.time_domain_args["offset"],
                   test_method_name="test_time_domain_impulse", report_func=report)

    def spectral_purity(self):
        func = dg_util.rpartial(functools.partial(self.pipeline, domain_name="freq"),
                                self.freq_domain_args["phase"], self.freq_domain_args["bin_offset"])

        def report(time_prod, freq_prod):
            d, t = time_prod["diff"], freq_prod["this"][1]
            return {"mean_diff": d["mean"], "total_diff": d["sum"], "max_spurious_power": t["max_spurious"],
                    "total_spurious_power": t["total_spurious"], "mean_spurious_power": t["mean_spurious"]}

        self._test(test_vector_func=func, test_vector_args=self.freq_domain_args["frequency"],
                   test_method_name="test_complex_sinusoid", report_func=report)

    def dispose(self):
        for f in self.files:
            f = getattr(f, "file_path", f)
            if os.path.exists(f):
                os.remove(f)
        self.files = []
        for f in glob.glob(os.path.join(self.output_dir, "channelized.*")):
            os.remove(f)

    def chop(self, input_dump_file, inverted_dump_file):       # purity.py:276-283
        input_dat = input_dump_file.data[self.total_sample_shift:, 0, :].flatten()
        inverted_dat = inverted_dump_file.data.flatten()
        return input_dat, inverted_dat

    def check(self, max_spurious_db=-60.0):
        """Hard pass/fail: every vector's maximum spurious power below the requirement."""
        worst = {}
        for name, rep in self.report.items():
            w = max(r["max_spurious_power"] for r in rep)
            worst[name] = w
            if w > max_spurious_db:
                raise AssertionError(f"{name}: max spurious power {w:.1f} dB exceeds {max_spurious_db} dB")
        return worst

    def finish(self):
        param_str = ".".join([f"fft_length-{self.input_fft_length}", f"deripple-{1 if self.deripple else 0}",
                              f"fft_window-{self.fft_window}", f"input_overlap-{self.input_overlap}"])
        path = os.path.join(self.products_dir, f"report.purity.{param_str}.json")
        os.makedirs(self.products_dir, exist_ok=True)
        with open(path, "w") as f:
            json.dump(self.report, f, cls=NumpyEncoder)
        return path


def main(argv=None):
    from ..data_gen.config import config_dir, load_config
    from .common import create_parser
    parsed = create_parser(description="B200 PFB inversion purity").parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if parsed.verbose else logging.INFO)
    cfg = load_config(parsed.sub_config_name)
    data_gen.set_config(cfg)
    test = TestPurity(n_test=parsed.n_test, os_factor=cfg["os_factor"], input_fft_length=cfg["input_fft_length"],
                      input_overlap=cfg["input_overlap"], fft_window=cfg.get("temporal_taper", "tukey"),
                      deripple=cfg.get("deripple", True), channels=cfg["channels"],
                      fir_filter_taps=cfg["fir_filter_taps"],
                      fir_filter_coeff_file_path=os.path.join(config_dir, cfg["fir_filter_coeff_file_path"]),
                      blocks=cfg["blocks"], backend=cfg.get("backend"), output_dir=parsed.output_dir,
                      save_output=parsed.save_output)
    if parsed.do_time:
        test.temporal_purity()
    if parsed.do_freq:
        test.spectral_purity()
    print(test.finish())
    return test


if __name__ == "__main__":
    main()
