"""Image I/O conventions and the saliency-mask front end of the reference (GAN_functions.py:24-208).

These are the steps *before* the accelerated path: file loading and the
opencv-contrib saliency pipeline.  They are kept as thin cv2 wrappers with the reference's formulas so
the entry points exist; `compute_saliency_map` needs `cv2.saliency` (opencv-contrib) and raises a clear
error when it is absent - callers then pass an explicit mask.
"""
from __future__ import annotations

import os

import numpy as np


def create_directories(directories):
    """GAN_functions.py:18-22."""
    for directory in directories:
        if not os.path.exists(directory):
            os.makedirs(directory)


def load_and_preprocess_image(image_path, target_size=(256, 256)):
    """GAN_functions.py:24-39: BGR file -> RGB, resize, (u8 - 127.5)/127.5."""
    import cv2
    img = cv2.imread(image_path)
    if img is None:
        raise ValueError(f"Could not load image: {image_path}")
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    img = cv2.resize(img, target_size)
    return (img.astype(np.float32) - 127.5) / 127.5


def save_image(img, path):
    """GAN_functions.py:41-50: ((img+1)*127.5).astype(uint8) (truncation), RGB -> BGR, imwrite."""
    import cv2
    out = ((np.asarray(img) + 1) * 127.5).astype(np.uint8)
    if out.ndim == 3 and out.shape[2] == 3:
        out = cv2.cvtColor(out, cv2.COLOR_RGB2BGR)
    cv2.imwrite(path, out)


def compute_saliency_map(image, method="spectral_residual"):
    """GAN_functions.py:52-121 (spectral residual 0.6 + fine grained 0.4 for 'combined')."""
    import cv2
    if not hasattr(cv2, "saliency"):
        raise RuntimeError("compute_saliency_map needs cv2.saliency (opencv-contrib), which is not installed; "
                           "pass a saliency mask explicitly (mask=...)")
    img_u8 = ((np.asarray(image) + 1) * 127.5).astype(np.uint8)
    bgr = cv2.cvtColor(img_u8, cv2.COLOR_RGB2BGR)

    def _run(algo):
        ok, sal = algo.computeSaliency(bgr)
        if not ok:
            raise RuntimeError("cv2.saliency failed")
        sal = sal.astype(np.float32)
        if sal.max() > 0:
            sal = sal / sal.max()
        return sal

    if method == "spectral_residual":
        return _run(cv2.saliency.StaticSaliencySpectralResidual_create())
    if method == "fine_grained":
        return _run(cv2.saliency.StaticSaliencyFineGrained_create())
    if method == "combined":
        a = _run(cv2.saliency.StaticSaliencySpectralResidual_create())
        b = _run(cv2.saliency.StaticSaliencyFineGrained_create())
        sal = 0.6 * a + 0.4 * b
        return sal / sal.max() if sal.max() > 0 else sal
    raise ValueError(f"unknown saliency method {method!r}")


def create_saliency_mask(saliency_map, threshold=None, smooth=True):
    """GAN_functions.py:159-208.  With smooth=True (the only mode the reference uses) the mask is
    bilateral(9,75,75) -> Gaussian 31x31 -> /max; the Otsu threshold is dead code in that mode."""
    import cv2
    sal = np.asarray(saliency_map, dtype=np.float32)
    if smooth:
        mask = cv2.bilateralFilter(sal, 9, 75, 75)
        mask = cv2.GaussianBlur(mask, (31, 31), 0)
        if mask.max() > 0:
            mask = mask / mask.max()
        return mask
    if threshold is None:
        u8 = (sal * 255).astype(np.uint8)
        thr, _ = cv2.threshold(u8, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        threshold = thr / 255.0
    return (sal > threshold).astype(np.float32)
