"""Drop-in for the inference-path names of the reference's GAN_functions.py, backed by libcic (B200).

`from GAN_functions import build_encoder, build_generator, ...` keeps working (GAN_test.py:14-20 of
the reference); each callable keeps the reference's signature.  See cic_b200.gan for the citations.
Training-only pieces (discriminator, SpectralNormalization) and matplotlib visualisation are out of
scope of the accelerated path and raise NotImplementedError.
"""
import cic_b200 as _cic
from cic_b200.gan import (  # noqa: F401
    AdaptiveQuantizationLayer, SelfAttention, build_adaptive_compression_model, build_encoder, build_generator,
    build_latent_saliency_model, build_rate_distortion_optimizer, compute_metrics, estimate_compression_ratio)
from cic_b200.saliency import (  # noqa: F401
    compute_saliency_map, create_directories, create_saliency_mask, enhance_saliency_map, load_and_preprocess_image, save_image)


def _out_of_scope(name, why):
    def _f(*a, **k):
        raise NotImplementedError(f"{name} is outside the accelerated inference path: {why}")
    _f.__name__ = name
    return _f


build_discriminator = _out_of_scope("build_discriminator", "training only (GAN_train.py:154)")
SpectralNormalization = _out_of_scope("SpectralNormalization", "dead code in the reference (never instantiated)")
visualize_results = _out_of_scope("visualize_results", "matplotlib plotting")
visualize_bit_allocation_by_bpp = _out_of_scope("visualize_bit_allocation_by_bpp", "matplotlib plotting")
