"""cic_b200: B200-native (sm_100a) inference hot path of Contextual-Image-Compression.

The directory is named `contextual-image-compression_b200`; import it as `cic_b200` (see cic_b200.py at
the repository root).  Importing this package loads libcic.so and fails loudly if it is not built.
"""
from . import _lib, runtime, weights, synth, models, ops, gan, autoencoder, saliency, dist, hdf5_lite, keras_h5  # noqa: F401
from .runtime import set_precision, get_precision  # noqa: F401
from ._lib import CicError  # noqa: F401

__all__ = ["_lib", "runtime", "weights", "synth", "models", "ops", "gan", "autoencoder", "saliency", "dist",
           "set_precision", "get_precision", "CicError"]
