"""Host-side mirror of train_autoencoder.py / test_autoencoder.py for the inference path."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from . import ops
from .models import AutoencoderModel
from .runtime import to_device_f32


def build_autoencoder(input_shape):
    """train_autoencoder.py:9-40: 7-conv U-Net with two skip connections, sigmoid output."""
    if len(input_shape) != 3:
        raise ValueError(f"input_shape must be (H, W, C), got {input_shape}")
    return AutoencoderModel(input_shape)


def load_images_from_folder(folder, target_size, with_paths=False):
    """train_autoencoder.py:42-56 / test_autoencoder.py:13-27: BGR, resized, /255."""
    import cv2
    images = []
    for path in glob.glob(os.path.join(folder, "*.jpg")):
        img = cv2.imread(path)
        if img is not None:
            img = cv2.resize(img, target_size).astype("float32") / 255.0
            images.append((img, path) if with_paths else img)
    return images if with_paths else np.array(images)


def _as_pair(image1, image2):
    a, b = np.asarray(image1), np.asarray(image2)
    if a.shape != b.shape:
        raise ValueError(f"Input images must have the same dimensions, got {a.shape} and {b.shape}")
    return a, b


def calculate_mse(image1, image2):
    """test_autoencoder.py:49-50.  On uint8 operands numpy wraps (a-b) and the square mod 256
    (SURVEY.md App. D.1); that behaviour is reproduced."""
    a, b = _as_pair(image1, image2)
    if a.dtype == np.uint8 and b.dtype == np.uint8:
        return float(ops.metrics_gray_u8(a, b)[0, 3].item())
    return float(ops.metrics_f32(a.astype(np.float32), b.astype(np.float32), signed_range=False)[0, 2].item())


def calculate_psnr(image1, image2):
    """test_autoencoder.py:52-57."""
    a, b = _as_pair(image1, image2)
    if a.dtype == np.uint8 and b.dtype == np.uint8:
        return float(ops.metrics_gray_u8(a, b)[0, 0].item())
    if a.dtype == np.uint8 or b.dtype == np.uint8:
        return float(ops.metrics_f32(a.astype(np.float32), b.astype(np.float32), False, 255.0)[0, 0].item())
    return float(ops.metrics_f32(a, b, signed_range=False, data_range=1.0)[0, 0].item())


def calculate_ssim(image1, image2):
    """test_autoencoder.py:59-66: SSIM of the BGR2GRAY images."""
    a, b = _as_pair(image1, image2)
    if a.dtype == np.uint8 and b.dtype == np.uint8:
        return float(ops.metrics_gray_u8(a, b)[0, 1].item())
    coef = np.array([0.114, 0.587, 0.299], dtype=np.float32)  # cv2 float BGR2GRAY
    ga = (a.astype(np.float32) * coef).sum(-1, keepdims=True).astype(np.float32)
    gb = (b.astype(np.float32) * coef).sum(-1, keepdims=True).astype(np.float32)
    return float(ops.metrics_f32(ga, gb, signed_range=False, data_range=1.0)[0, 1].item())


def evaluate_batch(model: AutoencoderModel, images, save_paths=None) -> dict:
    """The loop of test_autoencoder.py:83-108 for a whole batch in three launches-groups: predict,
    truncating uint8 cast of output and input, fused PSNR/SSIM/MSE.  images (B,H,W,3) float32 in [0,1].
    Returns per-image arrays 'mse' (the reference's wrapped value), 'true_mse', 'psnr', 'ssim'.
    save_paths: one file name per image = the reference's cv2.imwrite(compressed_path, compressed_img_uint8) (:90-93); .jpg names
    are encoded on the GPU (ops.jpeg_encode, the bytes OpenCV writes; the channel order is whatever the input had, as in the
    reference), other formats are written by OpenCV from the downloaded uint8 batch."""
    x = to_device_f32(images)
    y, y8 = model.forward_device([x], want_u8=True)
    x8 = ops.f32_to_u8_trunc(x, 255.0)                       # test_autoencoder.py:96
    m = ops.metrics_gray_u8(x8, y8).cpu().numpy()
    if save_paths is not None:
        import os
        if len(save_paths) != y8.shape[0]:
            raise ValueError(f"{len(save_paths)} paths for {y8.shape[0]} images")
        is_jpg = [os.path.splitext(p)[1].lower() in (".jpg", ".jpeg", ".jpe") for p in save_paths]
        if any(is_jpg):
            files = ops.jpeg_encode(y8)
            for ok, data, p in zip(is_jpg, files, save_paths):
                if ok:
                    with open(p, "wb") as f:
                        f.write(data)
        if not all(is_jpg):
            import cv2
            host = y8.cpu().numpy()
            for ok, img, p in zip(is_jpg, host, save_paths):
                if not ok:
                    cv2.imwrite(p, img)
    return {"psnr": m[:, 0], "ssim": m[:, 1], "true_mse": m[:, 2], "mse": m[:, 3], "compressed_u8": y8, "compressed": y}
