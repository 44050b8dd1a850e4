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
    """GAN_test.py:37-220: load the trained component models from `model_dir`.  Reads, in this order of preference,
      * the reference's own Keras checkpoints `<component>_final.h5`, or the latest `<component>_epoch_<n>.h5` when the final ones
        are missing (GAN_test.py:84-125) - through cic_b200.keras_h5 on the package's HDF5 reader (no h5py / TensorFlow needed);
      * `adaptive_weights.npz`, the flat Keras-layout checkpoint tools/convert_keras_h5.py writes;
    or takes `weights` ({sub_model: {name: array}}, e.g. cic_b200.weights.synthetic_adaptive) directly.  Builds the same eight-entry
    dict; with neither argument the models keep Keras-default init.  Raises ValueError("No models found! ...") like the reference
    (:219) when the directory holds neither."""
    if model_dir is not None and weights is None:
        import cic_b200
        npz = os.path.join(model_dir, "adaptive_weights.npz")
        if os.path.isdir(model_dir) and cic_b200.keras_h5.find_suffix(model_dir):
            weights = cic_b200.keras_h5.load_adaptive_dir(model_dir)
        elif os.path.exists(npz):
            weights = cic_b200.weights.load_npz(npz)
        else:
            raise ValueError("No models found! Please train the models first.")
        cic_b200.weights.check_adaptive(weights, IMG_SHAPE, BASE_LATENT_DIM)
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
