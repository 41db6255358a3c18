"""generate -> channelize -> synthesize chaining (python/data_gen/pipeline.py:13-86)."""
from __future__ import annotations

import os

__all__ = ["pipeline"]


def pipeline(test_vector_callback, channelize_callback, synthesize_callback, output_dir: str = None):
    def _pipeline(*args, **kwargs):
        test_vector_dada_file = test_vector_callback(*args, **kwargs, output_dir=output_dir)
        base = os.path.basename(test_vector_dada_file.file_path)
        channelized_dada_file = channelize_callback(test_vector_dada_file.file_path,
                                                    output_file_name="channelized." + base, output_dir=output_dir)
        synthesized_dada_file = synthesize_callback(channelized_dada_file.file_path,
                                                    output_file_name="synthesized." + base, output_dir=output_dir)
        return (test_vector_dada_file, channelized_dada_file, synthesized_dada_file)
    return _pipeline
