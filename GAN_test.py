"""Drop-in for the evaluation entry points of the reference's GAN_test.py (inference path only)."""
import os

import numpy as np

import cic_b200 as _cic
from cic_b200.gan import (  # noqa: F401
    BASE_LATENT_DIM, BPP_VALUES, HQ_LATENT_DIM, IMG_SHAPE, IMG_SIZE, bpp_accounting, compress_and_reconstruct,
    test_rate_control)
from GAN_functions import build_adaptive_compression_model, compute_metrics  # noqa: F401

TEST_DIR = "test_dataset"
RESULTS_DIR = "test_results"
MODEL_DIR = "models"


def load_models(model_dir=None, weights=None):
    """GAN_test.py:37-220 loads Keras .h5 checkpoints.  h5py is not available here, so `model_dir` is expected to hold
    `adaptive_weights.npz`, the flat Keras-layout checkpoint tools/convert_keras_h5.py writes from the reference's `*_final.h5`
    files; alternatively pass `weights` ({sub_model: {name: array}}, e.g. cic_b200.weights.synthetic_adaptive).  Builds the same
    eight-entry dict; with neither argument the models keep Keras-default init."""
    if model_dir is not None and weights is None:
        import os
        import cic_b200
        npz = os.path.join(model_dir, "adaptive_weights.npz")
        if os.path.exists(npz):                    # written by tools/convert_keras_h5.py from the reference's *_final.h5 files
            weights = cic_b200.weights.load_npz(npz)
            cic_b200.weights.check_adaptive(weights, IMG_SHAPE, BASE_LATENT_DIM)
        elif os.path.isdir(model_dir) and any(f.endswith(".h5") for f in os.listdir(model_dir)):
            raise NotImplementedError(f"{model_dir} holds Keras .h5 checkpoints: convert them once with tools/convert_keras_h5.py in the "
                                      "reference's TensorFlow environment (h5py is not available here) -> adaptive_weights.npz")
        else:
            raise FileNotFoundError(f"no adaptive_weights.npz (or .h5 checkpoints) in {model_dir}")
    models = build_adaptive_compression_model(IMG_SHAPE, BASE_LATENT_DIM, target_bpp=True)
    if weights is not None:
        models["adaptive_model"].set_weights_dict(weights)
    return models


def test_compression(test_images, file_names, original_sizes, models, masks=None):
    """GAN_test.py:342-455 without file output: every image x BPP_VALUES -> per-bpp metric lists + averages."""
    results_by_bpp = {bpp: {k: [] for k in ("psnr", "ssim", "mse", "compression_ratio", "actual_bpp", "hq_ratio")}
                      for bpp in BPP_VALUES}
    for i, img in enumerate(test_images):
        for bpp in BPP_VALUES:
            r = compress_and_reconstruct(img, models, target_bpp=bpp, mask=None if masks is None else masks[i])
            for k in ("psnr", "ssim", "mse"):
                results_by_bpp[bpp][k].append(r["metrics"][k])
            for k in ("compression_ratio", "actual_bpp", "hq_ratio"):
                results_by_bpp[bpp][k].append(r[k])
    avg_metrics = {bpp: {k: float(np.mean(v)) for k, v in res.items()} for bpp, res in results_by_bpp.items()}
    return {"results_by_bpp": results_by_bpp, "avg_metrics": avg_metrics}


test_compression.__test__ = False


def main():
    from cic_b200 import synth, weights
    print("\n===== B200 content-adaptive image compression: synthetic-input evaluation =====\n")
    models = load_models(weights=weights.synthetic_adaptive(IMG_SHAPE, BASE_LATENT_DIM))
    imgs = synth.to_signed_range(synth.synth_images_u8(4, *IMG_SIZE))
    masks = synth.synth_masks(4, *IMG_SIZE)[..., 0]
    names = [f"synthetic-{i}.png" for i in range(4)]
    res = test_compression(list(imgs), names, [0] * 4, models, masks=list(masks))
    for bpp, m in res["avg_metrics"].items():
        print(f"target {bpp} bpp: PSNR {m['psnr']:.2f} dB  SSIM {m['ssim']:.4f}  actual bpp {m['actual_bpp']:.4f}  "
              f"HQ ratio {m['hq_ratio'] * 100:.2f}%")
    rc = test_rate_control(models, list(imgs), names, masks=list(masks))
    for t, a, h in zip(rc["target_bpp"], rc["actual_bpp"], rc["hq_ratio"]):
        print(f"  target {t:.3f} -> actual {a:.4f} bpp, hq_ratio {h:.4f}")


if __name__ == "__main__":
    main()
