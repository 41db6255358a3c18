"""B200-native oversampled polyphase filterbank analysis + Golden FFT-based
inversion (the hot path of ska-telescope/ska-pst-dsp-model), behind the
reference's Python call surface.  CUDA only: importing ``ops`` without the
built extension raises."""
__version__ = "0.1.0"

from .rational import Rational  # noqa: F401
