"""Drop-in for the reference's test_autoencoder.py.

The reference is a module-level script (it loads `autoencoder_model.h5` and runs on import,
test_autoencoder.py:29-123).  Here the metric functions keep their names and the evaluation loop is
`main()`, which runs the same steps on a batch: predict, truncating uint8 cast, MSE/PSNR/SSIM.
"""
import numpy as np

import cic_b200 as _cic
from cic_b200.autoencoder import (  # noqa: F401
    build_autoencoder, calculate_mse, calculate_psnr, calculate_ssim, evaluate_batch, load_images_from_folder)

target_size = (128, 128)  # test_autoencoder.py:39


model_path = "autoencoder_model.h5"  # test_autoencoder.py:30


def load_model(path=model_path, input_shape=(128, 128, 3)):
    """test_autoencoder.py:30-34: the trained autoencoder from its Keras .h5 checkpoint (read with cic_b200.keras_h5)."""
    import os
    from cic_b200 import keras_h5
    if not os.path.exists(path):
        raise FileNotFoundError(f"Trained model not found: {path}")   # test_autoencoder.py:31-32
    model = build_autoencoder(input_shape)
    model.set_weights_dict(keras_h5.load_autoencoder(path))
    return model


def main(n_images=32, size=(256, 256), path=None):
    from cic_b200 import synth, weights
    if path is not None:
        model = load_model(path, (size[0], size[1], 3))
    else:
        model = build_autoencoder((size[0], size[1], 3))
        model.set_weights_dict(weights.synthetic_autoencoder())
    imgs = synth.to_unit_range(synth.synth_images_u8(n_images, size[0], size[1]))
    r = evaluate_batch(model, imgs)
    print("\n=== Overall Compression Performance ===")
    print(f"  - Average Mean Squared Error (MSE): {np.mean(r['mse']):.4f}")
    print(f"  - Average Peak Signal-to-Noise Ratio (PSNR): {np.mean(r['psnr']):.2f} dB")
    print(f"  - Average Structural Similarity Index (SSIM): {np.mean(r['ssim']):.4f}")
    return r


if __name__ == "__main__":
    main()
