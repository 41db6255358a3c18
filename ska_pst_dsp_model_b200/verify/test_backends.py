"""Cross-backend channelizer comparison (python/verify/test_backends.py:28-119): the same
complex sinusoid is channelized by every backend and compared with
``isclose(atol=rtol=1e-4)``.  Unlike the reference this raises on mismatch."""
from __future__ import annotations

import logging
import os
import tempfile

import numpy as np

from ..data_gen import channelize, generate_test_vector
from ..rational import Rational

module_logger = logging.getLogger(__name__)


class TestBackends:
    __test__ = False
    thresh = 1e-4

    def __init__(self, os_factor, input_fft_length, channels, fir_filter_coeff_file_path, blocks, use_padded=False,
                 backends=("b200", "python"), output_dir=None):
        self.os_factor = Rational.from_str(os_factor)
        self.channels = channels
        self.n_samples = self.os_factor.normalize(input_fft_length) * channels * blocks      # :52-54
        self.fir_filter_path = fir_filter_coeff_file_path
        self.use_padded = use_padded
        self.backends = list(backends)
        self.output_dir = output_dir or tempfile.mkdtemp(prefix="pfb_backends_")
        self.results = {}

    def test_channelizer_backends(self):
        input_file = generate_test_vector([0.001], [np.pi / 4], 0.0, n_bins=self.n_samples, domain_name="freq",
                                          backend="python", output_dir=self.output_dir)
        channelized = []
        for b in self.backends:
            file_name = f"channelized.{b}.{os.path.basename(input_file.file_path)}"
            channelized.append(channelize(backend=b, use_padded=self.use_padded)(
                input_file.file_path, channels=self.channels, os_factor_str=str(self.os_factor),
                fir_filter_path=self.fir_filter_path, output_dir=self.output_dir, output_file_name=file_name))
        ref = channelized[0].data.flatten()
        for b, d in zip(self.backends[1:], channelized[1:]):
            other = d.data.flatten()
            if other.shape != ref.shape:
                raise AssertionError(f"backend {b}: shape {other.shape} != {ref.shape}")
            close = np.isclose(other, ref, atol=self.thresh, rtol=self.thresh)
            self.results[b] = {"mean": float(np.mean(close)), "sum": int(np.sum(close)),
                               "max_diff": float(np.abs(other - ref).max())}
            if self.results[b]["mean"] != 1.0:
                raise AssertionError(f"Channelizers not equal for backend {b}: {self.results[b]}")
        return self.results
