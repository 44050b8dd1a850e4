"""Freeze outputs of the CPU oracle on seeded inputs into tests/golden/oracle_small.npz.

The reference itself cannot be imported in this container (TensorFlow / scikit-image absent), so these
vectors pin the *oracle* against drift and give the GPU tests a fixture that does not need the oracle's
heavy graph at run time.  Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import importlib

synth = importlib.import_module("contextual-image-compression_b200.synth")
W = importlib.import_module("contextual-image-compression_b200.weights")
from oracle import graphs, metrics  # noqa: E402

torch.set_num_threads(1)  # deterministic reduction order
IMG = (64, 64, 3)
BASE = 32


def main():
    out = {}
    # autoencoder, 2 x 32 x 48 images
    aw = W.synthetic_autoencoder(seed=42)
    x = synth.to_unit_range(synth.synth_images_u8(2, 32, 48, seed=43))
    y = graphs.autoencoder_forward(aw, x)
    out["ae_y"] = y
    y8, x8 = graphs.autoencoder_output_u8(y), (x * 255).astype(np.uint8)
    out["ae_metrics"] = np.array([[metrics.ae_calculate_mse(a, b), metrics.ae_true_mse(a, b), metrics.ae_calculate_psnr(a, b),
                                   metrics.ae_calculate_ssim(a, b)] for a, b in zip(x8, y8)])
    # adaptive codec, 3 tiles of 64x64, base latent 32
    ws = W.synthetic_adaptive(IMG, BASE, seed=42)
    img = synth.to_signed_range(synth.synth_images_u8(3, 64, 64, seed=44))
    mask = synth.synth_masks(3, 64, 64, seed=44)
    bpp = np.array([[0.1], [1.0], [2.0]], np.float32)
    outs, ex = graphs.adaptive_forward(ws, img, mask, bpp, return_extras=True)
    for name, v in zip(("blended", "hq_q", "lq_q", "rd", "dt"), outs):
        out["ad_" + name] = v
    for k in ("hq_latent", "lq_latent", "hq_sym", "lq_sym", "hq_scale", "lq_scale", "sal_hq", "sal_lq"):
        out["ad_" + k] = ex[k]
    out["ad_metrics"] = np.array([[m["psnr"], m["ssim"], float(m["mse"])] for m in
                                  (metrics.compute_metrics(a, b) for a, b in zip(img, outs[0]))])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
