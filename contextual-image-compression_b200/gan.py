"""Host-side mirror of the reference's GAN codec interface (GAN_functions.py / GAN_test.py).

Same names, argument meaning and error behaviour as the reference for the hot path; the
arithmetic runs in libcic.  File I/O, saliency extraction (opencv-contrib) and plotting are
outside the accelerated path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops, runtime, weights as W
from .models import (AdaptiveCompressionModel, EncoderModel, GeneratorModel, LatentSaliencyModel, Model,
                     RDOptimizerModel)
from .runtime import to_device_f32

# configuration constants of GAN_test.py:23-29
IMG_SIZE = (256, 256)
IMG_SHAPE = (IMG_SIZE[0], IMG_SIZE[1], 3)
BASE_LATENT_DIM = 512
HQ_LATENT_DIM = BASE_LATENT_DIM * 2
BPP_VALUES = [0.1, 1.0, 2.0]


# ---- builders (GAN_functions.py) ---------------------------------------------------------------
def build_latent_saliency_model(latent_dim, name="latent_saliency_module"):
    """GAN_functions.py:210-234."""
    return LatentSaliencyModel(latent_dim, name=name)


def build_generator(latent_dim, img_shape, name="generator"):
    """GAN_functions.py:236-278.  The reference hard-wires 256x256 skip shapes (:242-244); here the
    skip shapes follow img_shape (H, W divisible by 16) and coincide with the reference at 256."""
    return GeneratorModel(latent_dim, img_shape, name=name)


def build_encoder(img_shape, latent_dim, name="encoder", add_attention=True):
    """GAN_functions.py:280-331 -> model returning [latent, x1, x2, x3]."""
    return EncoderModel(img_shape, latent_dim, name=name, add_attention=add_attention)


def build_rate_distortion_optimizer(img_shape, latent_dims, name="rd_optimizer"):
    """GAN_functions.py:495-557 (latent_dims is unused by the reference too)."""
    return RDOptimizerModel(img_shape, latent_dims, name=name)


def build_adaptive_compression_model(img_shape, base_latent_dim, target_bpp=None):
    """GAN_functions.py:559-722 -> dict with the reference's eight keys.

    The reference's `target_bpp=None` branch builds a disconnected Keras graph;
    the model here always takes [image, saliency, target_bpp].
    """
    comps: Dict[str, Model] = {
        "hq_encoder": build_encoder(img_shape, base_latent_dim * 2, name="hq_encoder", add_attention=True),
        "hq_generator": build_generator(base_latent_dim * 2, img_shape, name="hq_generator"),
        "lq_encoder": build_encoder(img_shape, base_latent_dim, name="lq_encoder", add_attention=False),
        "lq_generator": build_generator(base_latent_dim, img_shape, name="lq_generator"),
        "latent_saliency_hq": build_latent_saliency_model(base_latent_dim * 2, name="hq_latent_saliency"),
        "latent_saliency_lq": build_latent_saliency_model(base_latent_dim, name="lq_latent_saliency"),
        "rd_optimizer": build_rate_distortion_optimizer(
            img_shape, {"hq": base_latent_dim * 2, "lq": base_latent_dim}, name="rd_optimizer"),
    }
    adaptive = AdaptiveCompressionModel(img_shape, base_latent_dim, comps)
    out = {"adaptive_model": adaptive}
    out.update(comps)
    return out


# ---- layers --------------------------------------------------------------------------------------
class SelfAttention:
    """GAN_functions.py:333-374: SAGAN-style attention, no 1/sqrt(d) scaling, gamma initialised to 0."""

    def __init__(self, channels, seed: int = 0, **kwargs):
        self.channels = int(channels)
        rng = np.random.Generator(np.random.PCG64(seed))
        c = self.channels
        self.query_kernel = W._glorot(rng, (1, 1, c, c // 8))
        self.key_kernel = W._glorot(rng, (1, 1, c, c // 8))
        self.value_kernel = W._glorot(rng, (1, 1, c, c))
        self.query_bias = np.zeros(c // 8, np.float32)
        self.key_bias = np.zeros(c // 8, np.float32)
        self.value_bias = np.zeros(c, np.float32)
        self.gamma = np.zeros(1, np.float32)

    def __call__(self, inputs):
        y = ops.self_attention(inputs, self.query_kernel, self.query_bias, self.key_kernel, self.key_bias,
                               self.value_kernel, self.value_bias, float(self.gamma[0]))
        return y.cpu()

    call = __call__

    def get_config(self):
        return {"channels": self.channels}


class AdaptiveQuantizationLayer:
    """GAN_functions.py:429-446: quantized = round(latent*scale)/scale, scale = exp(3*qs*(1-sal))."""

    def __init__(self, **kwargs):
        self.name = kwargs.get("name", "adaptive_quantization")

    def __call__(self, inputs):
        latent, saliency_score, quant_strength = inputs
        return ops.quantize_latent(latent, saliency_score, quant_strength)["deq"].cpu()

    call = __call__


# ---- evaluation (GAN_functions.py:724-823, GAN_test.py:265-340, 532-645) -----------------------------
def compute_metrics(original_img, compressed_img):
    """GAN_functions.py:724-759: images in [-1,1]; returns {'psnr','ssim','mse'}."""
    a = to_device_f32(original_img)
    b = to_device_f32(compressed_img)
    if a.dim() == 2:
        a, b = a.unsqueeze(-1), b.unsqueeze(-1)
    m = ops.metrics_f32(a, b, signed_range=True, data_range=1.0)[0].cpu().numpy()
    return {"psnr": float(m[0]), "ssim": float(m[1]), "mse": np.float32(m[2])}


def compute_metrics_batch(original, compressed) -> np.ndarray:
    """(B,3) float64 [psnr, ssim, mse] for a batch of [-1,1] images - one kernel launch."""
    return ops.metrics_f32(original, compressed, signed_range=True, data_range=1.0)[:, :3].cpu().numpy()


def estimate_compression_ratio(original_size, latent_size):
    """GAN_functions.py:809-823."""
    compression_ratio = original_size / latent_size
    percentage_reduction = (1 - (latent_size / original_size)) * 100
    return compression_ratio, percentage_reduction


def bpp_accounting(hq_ratio, img_size=IMG_SIZE, base_latent_dim=BASE_LATENT_DIM):
    """GAN_test.py:310-325 (and :573-582): nominal bits from the HQ/LQ area split."""
    lq_ratio = 1.0 - hq_ratio
    hq_bits = hq_ratio * (base_latent_dim * 2) * 32
    lq_bits = lq_ratio * base_latent_dim * 32
    total_bits = hq_bits + lq_bits
    original_bits = img_size[0] * img_size[1] * 3 * 8
    return {"hq_ratio": hq_ratio, "lq_ratio": lq_ratio, "compression_ratio": original_bits / total_bits,
            "actual_bpp": total_bits / (img_size[0] * img_size[1])}


def _mask_for(img, mask, saliency_map=None):
    """The mask input of the model: an explicit mask, or create_saliency_mask(saliency_map, smooth=True) on the GPU when the caller
    has a saliency map, or the reference's whole front end (compute_saliency_map(img, 'combined') -> create_saliency_mask) as
    one device-side chain."""
    if mask is not None:
        return np.asarray(mask, dtype=np.float32).reshape(img.shape[0], img.shape[1])
    if saliency_map is None:
        return ops.saliency_mask_from_image(img, method="combined").cpu().numpy()         # GAN_test.py:279-280
    return ops.saliency_mask_smooth(np.asarray(saliency_map, np.float32)).cpu().numpy()   # GAN_test.py:280


def compress_and_reconstruct(img, models, target_bpp=1.0, mask=None, saliency_map=None):
    """GAN_test.py:265-340.  `mask` (H,W) in [0,1] replaces the reference's saliency front end (GAN_test.py:279-280);
    `saliency_map` replaces only its opencv-contrib half (compute_saliency_map) - the mask is then made from it on the GPU;
    with neither, the whole front end runs on the GPU (csrc/saliency_map.cu + saliency_mask.cu)."""
    img = np.asarray(img, dtype=np.float32)
    mask = _mask_for(img, mask, saliency_map)
    img_batch = np.expand_dims(img, axis=0)
    mask_batch = np.expand_dims(np.expand_dims(mask, axis=-1), axis=0)
    target_bpp_batch = np.array([[target_bpp]], dtype=np.float32)
    adaptive_model = models["adaptive_model"]
    compressed_output, hq_latent, lq_latent, rd_params, bit_allocation = adaptive_model.predict(
        [img_batch, mask_batch, target_bpp_batch], verbose=0)
    compressed_img = compressed_output[0]
    bit_alloc_map = bit_allocation[0]
    quality_metrics = compute_metrics(img, compressed_img)
    hq_ratio = np.mean(bit_alloc_map)
    acc = bpp_accounting(hq_ratio, adaptive_model.img_shape[:2], adaptive_model.base_latent_dim)
    return {
        "saliency_map": mask,
        "compressed_img": compressed_img,
        "hq_latent": hq_latent[0] if hq_latent.shape[0] == 1 else hq_latent,
        "lq_latent": lq_latent[0] if lq_latent.shape[0] == 1 else lq_latent,
        "rd_params": rd_params[0] if rd_params.shape[0] == 1 else rd_params,
        "bit_allocation": bit_alloc_map,
        "metrics": quality_metrics,
        "compression_ratio": acc["compression_ratio"],
        "actual_bpp": acc["actual_bpp"],
        "target_bpp": target_bpp,
        "hq_ratio": hq_ratio,
        "lq_ratio": acc["lq_ratio"],
    }


def test_rate_control(models, test_images, file_names, masks=None, full_model=False):
    """GAN_test.py:532-645 without the plots: first 4 images x np.linspace(0.1, 2.0, 10).

    hq_ratio depends only on the mask and the target bpp (GAN_functions.py:631-657), so by default
    one `cic_hq_ratio_sweep` pass per image replaces the reference's 10 full-model predictions;
    `full_model=True` runs the whole model per level like the reference and gives the same numbers.
    """
    test_bpps = np.linspace(0.1, 2.0, 10)
    adaptive_model = models["adaptive_model"]
    base = adaptive_model.base_latent_dim
    results = {"target_bpp": [], "actual_bpp": [], "hq_ratio": [], "image": []}
    n = min(4, len(test_images))
    for i in range(n):
        img = np.asarray(test_images[i], dtype=np.float32)
        mask = _mask_for(img, None if masks is None else masks[i])
        size = adaptive_model.img_shape[:2]  # bits are nominal per model tile (GAN_test.py:318-325)
        if full_model:
            ratios = []
            for bpp in test_bpps:
                outs = adaptive_model.predict([img[None], mask[None, :, :, None], np.array([[bpp]], np.float32)], verbose=0)
                ratios.append(float(np.mean(outs[4])))
        else:
            ratios = ops.hq_ratio_sweep(mask[None, :, :, None], test_bpps.astype(np.float32))[0].cpu().numpy().tolist()
        for bpp, hq_ratio in zip(test_bpps, ratios):
            acc = bpp_accounting(hq_ratio, size, base)
            results["target_bpp"].append(bpp)
            results["actual_bpp"].append(acc["actual_bpp"])
            results["hq_ratio"].append(hq_ratio)
            results["image"].append(file_names[i])
    return {"target_bpp": results["target_bpp"], "actual_bpp": results["actual_bpp"], "hq_ratio": results["hq_ratio"]}


test_rate_control.__test__ = False  # a reference entry point, not a pytest test
