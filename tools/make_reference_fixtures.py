#!/usr/bin/env python
"""Generate golden vectors FROM THE REFERENCE ITSELF -> tests/golden/reference_*.npz  (the route from "parity unpinned" to pinned).

Run it ONCE in the reference's own environment (TensorFlow / Keras 3, scikit-image, opencv; numpy) from this repository's root:

    python tools/make_reference_fixtures.py --reference /path/to/Contextual-Image-Compression

It imports the reference's `GAN_functions.py` and `train_autoencoder.py` unmodified, builds
`build_adaptive_compression_model((256,256,3), 512, target_bpp=True)` (GAN_functions.py:559) and `build_autoencoder`
(train_autoencoder.py:9), assigns the seeded synthetic weights this repository's tests use (weights.py, Keras layouts, so nothing is
transposed; the assignment walks `model.layers` in creation order - the inverse of tools/convert_keras_h5.py's `map_layers`), feeds
the seeded synthetic inputs (synth.py: integer-hash images, Gaussian-blob masks) and writes inputs' seeds + the reference's outputs:

  reference_adaptive.npz   adaptive_model.predict([img, mask, bpp]) -> blended, hq_latent_q, lq_latent_q, rd_params, dt for 2 tiles at
                           bpp 0.1 / 1.0; hq_encoder / lq_encoder latents (pre-quantisation); compute_metrics(img, blended) (:724)
  reference_autoencoder.npz  build_autoencoder((64,96,3)).predict(x) for 2 images
  reference_metrics.npz    skimage PSNR / SSIM on the uint8 autoencoder convention (test_autoencoder.py:52-66 formulas)
  reference_saliency.npz   compute_saliency_map (all three methods) and create_saliency_mask(smooth=True) on 3 synthetic 256x256 images
                           (only where cv2.saliency = opencv-contrib is installed)

tests/test_reference_fixtures.py consumes the files when they exist (CPU test: oracle vs fixture; GPU test: CUDA path vs fixture)
and skips with a message when they do not.

THIS SCRIPT CANNOT BE EXECUTED IN THIS REPOSITORY'S CONTAINER (no TensorFlow / scikit-image, no network).  Its pure-Python part
(`assign_layers`) is exercised on duck-typed layers by tests/test_host_logic.py.
"""
from __future__ import annotations

import argparse
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
IMG_SHAPE, BASE = (256, 256, 3), 512
SEED_W, SEED_X = 42, 45            # weights / inputs (the same seeds tests/test_gpu_models.py::test_adaptive_reference_size uses)


def pure(name):
    spec = importlib.util.spec_from_file_location(f"cic_pure_{name}", os.path.join(ROOT, "contextual-image-compression_b200", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cls(layer) -> str:
    return type(layer).__name__


def assign_layers(layers, kind: str, w: dict) -> int:
    """Set the weights of one reference sub-model from a {name: array} dict in this repository's naming (weights.py).  Layers are
    taken in creation order per class, exactly like tools/convert_keras_h5.py::map_layers reads them.  Returns #tensors set."""
    convs = [l for l in layers if _cls(l) == "Conv2D"]
    deconvs = [l for l in layers if _cls(l) == "Conv2DTranspose"]
    denses = [l for l in layers if _cls(l) == "Dense"]
    bns = [l for l in layers if _cls(l) == "BatchNormalization"]
    attn = [l for l in layers if _cls(l) == "SelfAttention"]
    n = 0

    def put(layer, prefix, names=("kernel", "bias")):
        nonlocal n
        layer.set_weights([np.asarray(w[f"{prefix}/{k}"], np.float32) for k in names])
        n += len(names)

    bn_names = ("gamma", "beta", "moving_mean", "moving_variance")
    if kind == "encoder":
        assert len(convs) == 4 and len(bns) == 3 and len(denses) == 1, (len(convs), len(bns), len(denses))
        for i, l in enumerate(convs, start=1):
            put(l, f"conv{i}")
        for i, l in enumerate(bns, start=2):
            put(l, f"bn{i}", bn_names)
        put(denses[0], "dense")
        if attn:
            a = attn[0]
            for nm, sub in (("query", a.query_conv), ("key", a.key_conv), ("value", a.value_conv)):
                put(sub, f"attn/{nm}")
            a.gamma.assign(np.asarray(w["attn/gamma"], np.float32).reshape(1))
            n += 1
    elif kind == "generator":
        assert len(denses) == 1 and len(bns) == 5 and len(deconvs) == 4 and len(convs) == 1
        put(denses[0], "dense")
        for i, l in enumerate(bns):
            put(l, f"bn{i}", bn_names)
        for i, l in enumerate(deconvs, start=1):
            put(l, f"deconv{i}")
        put(convs[0], "conv_out")
    elif kind == "latent_saliency":
        assert len(denses) == 3
        for i, l in enumerate(denses, start=1):
            put(l, f"dense{i}")
    elif kind == "rd_optimizer":
        assert len(convs) == 2 and len(denses) == 2
        for i, l in enumerate(convs, start=1):
            put(l, f"conv{i}")
        for i, l in enumerate(denses, start=1):
            put(l, f"dense{i}")
    elif kind == "autoencoder":                   # train_autoencoder.py:14-35: seven Conv2D in creation order
        assert len(convs) == 7
        for l, nm in zip(convs, ("conv1", "conv2", "conv3", "conv_x2", "conv5", "conv_x1", "conv_out")):
            put(l, nm)
    else:
        raise ValueError(kind)
    return n


SUB_MODELS = (("hq_encoder", "encoder"), ("hq_generator", "generator"), ("lq_encoder", "encoder"), ("lq_generator", "generator"),
              ("latent_saliency_hq", "latent_saliency"), ("latent_saliency_lq", "latent_saliency"), ("rd_optimizer", "rd_optimizer"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of hassanrizwank/Contextual-Image-Compression")
    args = ap.parse_args()
    sys.path.insert(0, os.path.abspath(args.reference))
    import GAN_functions as ref                                   # noqa: PLC0415  the unmodified reference
    import train_autoencoder as ref_ae                            # noqa: PLC0415
    from skimage.metrics import peak_signal_noise_ratio, structural_similarity  # noqa: PLC0415
    import cv2                                                    # noqa: PLC0415
    synth, W = pure("synth"), pure("weights")
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- adaptive codec (GAN_functions.py:559-722) ------------------------------------------------------------------------------
    ws = W.synthetic_adaptive(IMG_SHAPE, BASE, seed=SEED_W)
    models = ref.build_adaptive_compression_model(IMG_SHAPE, BASE, target_bpp=True)
    for sub, kind in SUB_MODELS:
        assign_layers(list(models[sub].layers), kind, ws[sub])
    img = synth.to_signed_range(synth.synth_images_u8(2, 256, 256, seed=SEED_X))
    mask = synth.synth_masks(2, 256, 256, seed=SEED_X)
    bpp = np.array([[0.1], [1.0]], np.float32)
    blended, hq_q, lq_q, rd, dt = models["adaptive_model"].predict([img, mask, bpp], verbose=0)
    hq_lat = models["hq_encoder"].predict(img, verbose=0)[0]
    lq_lat = models["lq_encoder"].predict(img, verbose=0)[0]
    mets = [ref.compute_metrics(img[i], blended[i]) for i in range(2)]
    np.savez_compressed(os.path.join(GOLDEN, "reference_adaptive.npz"),
                        seed_weights=SEED_W, seed_inputs=SEED_X, bpp=bpp, blended=blended, hq_latent_q=hq_q, lq_latent_q=lq_q,
                        rd_params=rd, dt=dt, hq_latent=hq_lat, lq_latent=lq_lat,
                        metrics=np.array([[m["psnr"], m["ssim"], float(m["mse"])] for m in mets], np.float64))

    # ---- autoencoder (train_autoencoder.py:9-40) -----------------------------------------------------------------------------------
    aw = W.synthetic_autoencoder(seed=SEED_W)
    ae = ref_ae.build_autoencoder((64, 96, 3))
    assign_layers(list(ae.layers), "autoencoder", aw)
    x = synth.to_unit_range(synth.synth_images_u8(2, 64, 96, seed=43))
    y = ae.predict(x, verbose=0)
    np.savez_compressed(os.path.join(GOLDEN, "reference_autoencoder.npz"), seed_weights=SEED_W, seed_inputs=43, y=y)

    # ---- metric call sites of test_autoencoder.py:49-66 on uint8 images ------------------------------------------------------------
    a8 = (x * 255).astype(np.uint8)
    b8 = (y * 255).astype(np.uint8)
    rows = []
    for i in range(2):
        mse = float(np.mean((a8[i] - b8[i]) ** 2))                                       # :49-50 (uint8 arithmetic, wraps)
        psnr = float(peak_signal_noise_ratio(a8[i], b8[i], data_range=255))              # :57
        ga, gb = cv2.cvtColor(a8[i], cv2.COLOR_BGR2GRAY), cv2.cvtColor(b8[i], cv2.COLOR_BGR2GRAY)   # :64-65
        rows.append([mse, psnr, float(structural_similarity(ga, gb, data_range=255))])   # :66
    np.savez_compressed(os.path.join(GOLDEN, "reference_metrics.npz"), a8=a8, b8=b8, rows=np.array(rows, np.float64))
    # ---- saliency front end (GAN_functions.py:52-121, :159-208; needs opencv-contrib: cv2.saliency) ------------------------------
    if hasattr(cv2, "saliency"):
        simg = synth.to_signed_range(synth.synth_images_u8(3, 256, 256, seed=47))
        sal = {m: np.stack([ref.compute_saliency_map(simg[i], method=m) for i in range(3)]).astype(np.float32)
               for m in ("spectral_residual", "fine_grained", "combined")}
        masks = np.stack([ref.create_saliency_mask(sal["combined"][i], smooth=True) for i in range(3)]).astype(np.float32)
        np.savez_compressed(os.path.join(GOLDEN, "reference_saliency.npz"), seed_inputs=47, masks=masks, **sal)
    else:
        print("cv2.saliency (opencv-contrib) is absent: reference_saliency.npz not written")
    import tensorflow as tf                                       # noqa: PLC0415
    import skimage                                                # noqa: PLC0415
    with open(os.path.join(GOLDEN, "reference_versions.txt"), "w") as f:
        f.write(f"tensorflow {tf.__version__}\nscikit-image {skimage.__version__}\nopencv {cv2.__version__}\nnumpy {np.__version__}\n")
    print("wrote tests/golden/reference_{adaptive,autoencoder,metrics}.npz")


if __name__ == "__main__":
    main()
