"""Configuration loader (python/data_gen/config.py:35-41).  The reference's default
sub-configuration name ``"test"`` does not exist in its shipped JSON (KeyError at
import); here ``"test"`` is defined as the 8-channel 8/7 PST-taps case of
BASELINE configs[0]."""
from __future__ import annotations

import json
import os

from . import util

__all__ = ["load_config", "config", "config_dir", "build_dir", "set_config"]

cur_dir = util.curdir(__file__)
config_dir = os.path.join(util.updir(cur_dir, 1), "config")
build_dir = os.path.join(util.updir(cur_dir, 2), "build")
test_config_file_name = "test.config.json"
test_config_file_path = os.path.join(config_dir, test_config_file_name)


def load_config(sub_config_name: str = "test"):
    if not os.path.exists(test_config_file_path):
        raise RuntimeError(f"Cannot find {test_config_file_name} in {config_dir}")
    with open(test_config_file_path, "r") as f:
        cfg = json.load(f)
    return cfg[sub_config_name]


config = load_config()


def set_config(sub_config_name_or_dict):
    """Switch the module-level defaults used by channelize/synthesize/generate_test_vector."""
    new = load_config(sub_config_name_or_dict) if isinstance(sub_config_name_or_dict, str) else dict(sub_config_name_or_dict)
    config.clear()
    config.update(new)
    return config
