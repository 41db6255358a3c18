"""The reference's ``data_gen`` call surface (python/data_gen/__init__.py) on the B200 backend."""
from .channelize import channelize
from .config import config, config_dir, load_config, set_config
from .dispose import dispose
from .generate_test_vector import complex_sinusoid, generate_test_vector, time_domain_impulse
from .pipeline import pipeline
from .synthesize import synthesize

__all__ = ["generate_test_vector", "complex_sinusoid", "time_domain_impulse", "channelize", "synthesize",
           "pipeline", "dispose", "config", "config_dir", "load_config", "set_config"]
