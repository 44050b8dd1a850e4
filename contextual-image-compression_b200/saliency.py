"""Image I/O conventions and the saliency-mask front end of the reference (GAN_functions.py:24-208).

These are the steps *before* the accelerated path: file loading and the opencv-contrib saliency
pipeline.  They follow the reference operation for operation so that masks (and therefore dt, hq_ratio, actual_bpp and the
blended image) are the reference's:

  * `compute_saliency_map` needs `cv2.saliency` (opencv-contrib, not installed in this image); it raises a clear error when the
    module is absent - callers then pass an explicit mask - and follows the reference's fall-backs when a method *fails*.
  * `create_saliency_mask(smooth=True)` - the only mode the reference uses (GAN_test.py:280,553) - is plain OpenCV (bilateral
    9/75/75, Gaussian 31x31, / max); `ops.saliency_mask_smooth` is the same arithmetic on the GPU (§8 f2).
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
    """GAN_functions.py:41-50: ((img+1)*127.5).astype(uint8) (truncation), RGB -> BGR, imwrite.  A 3-channel image saved as
    .jpg / .jpeg is encoded on the GPU (ops.jpeg_encode: the same bytes cv2.imwrite produces, with the RGB -> BGR swap folded into
    the kernel); other formats and grey images go through OpenCV."""
    out = ((np.asarray(img) + 1) * 127.5).astype(np.uint8)
    if out.ndim == 3 and out.shape[2] == 3 and os.path.splitext(path)[1].lower() in (".jpg", ".jpeg", ".jpe"):
        from . import ops
        with open(path, "wb") as f:
            f.write(ops.jpeg_encode(out, rgb=True))
        return
    import cv2
    if out.ndim == 3 and out.shape[2] == 3:
        out = cv2.cvtColor(out, cv2.COLOR_RGB2BGR)
    cv2.imwrite(path, out)


def save_images_u8(images_u8, paths, rgb=True, quality=95):
    """Batch form of save_image for images that are already uint8 on the device (predict_phased / predict_stream with u8_io=True
    give ((x + 1) * 127.5).astype(uint8) RGB): one encoder call, only the compressed bytes cross PCIe.  .jpg paths only."""
    from . import ops
    files = ops.jpeg_encode(images_u8, quality=quality, rgb=rgb)
    if isinstance(files, bytes):
        files = [files]
    if len(files) != len(paths):
        raise ValueError(f"{len(files)} images for {len(paths)} paths")
    for data, path in zip(files, paths):
        if os.path.splitext(path)[1].lower() not in (".jpg", ".jpeg", ".jpe"):
            raise ValueError(f"save_images_u8 writes JPEG files, got '{path}'")
        with open(path, "wb") as f:
            f.write(data)


def _saliency_module():
    import cv2
    mod = getattr(cv2, "saliency", None)
    if mod is None:
        raise RuntimeError("compute_saliency_map needs cv2.saliency (opencv-contrib), which is not installed; "
                           "pass a saliency mask explicitly (mask=...)")
    return mod


def _to_cv_bgr_u8(image):
    """GAN_functions.py:63-71: float32 images with max <= 1 are taken as [-1, 1] and mapped to [0, 255] (truncation); anything
    else is cast to uint8 as is.  3-channel images are RGB -> BGR."""
    import cv2
    image = np.asarray(image)
    if image.dtype == np.float32 and np.max(image) <= 1.0:
        image_cv = ((image + 1) * 127.5).astype(np.uint8)
    else:
        image_cv = image.astype(np.uint8)
    if image_cv.ndim == 3 and image_cv.shape[2] == 3:
        image_cv = cv2.cvtColor(image_cv, cv2.COLOR_RGB2BGR)
    return image_cv


def compute_saliency_map(image, method="spectral_residual"):
    """GAN_functions.py:52-121.

    'combined' mixes the RAW spectral-residual and fine-grained maps (0.6 / 0.4, :95) and normalises only the sum (:98-99) - the
    maps keep their native ranges in the mix.  A failed method falls back to the surviving map (returned un-normalised, as the
    reference does, :84-88) or to a uniform map of ones (:89-91, :112-115)."""
    sal = _saliency_module()
    image_cv = _to_cv_bgr_u8(image)
    uniform = np.ones(image_cv.shape[:2], dtype=np.float32)
    if method == "combined":
        ok_s, spectral = sal.StaticSaliencySpectralResidual_create().computeSaliency(image_cv)
        ok_f, fine = sal.StaticSaliencyFineGrained_create().computeSaliency(image_cv)
        if not (ok_s and ok_f):
            print("Warning: One or more saliency methods failed. Using available method.")
            if ok_s:
                return spectral
            if ok_f:
                return fine
            print("All saliency methods failed. Returning uniform saliency.")
            return uniform
        mix = 0.6 * spectral + 0.4 * fine
        peak = mix.max()
        return mix / peak if peak > 0 else mix
    if method == "spectral_residual":
        algo = sal.StaticSaliencySpectralResidual_create()
    elif method == "fine_grained":
        algo = sal.StaticSaliencyFineGrained_create()
    else:
        raise ValueError(f"Unsupported saliency method: {method}")
    ok, out = algo.computeSaliency(image_cv)
    if not ok:
        print(f"Failed to compute saliency using {method} method.")
        return uniform
    peak = out.max()
    return out / peak if peak > 0 else out


def adaptive_threshold(saliency_map):
    """The threshold create_saliency_mask derives when none is given (GAN_functions.py:172-194): min(Otsu on the uint8 map, the
    lower edge of the first of 50 histogram bins whose cumulative share exceeds 0.7), clamped to [0.05, 0.5]."""
    import cv2
    sal = np.asarray(saliency_map)
    u8 = (sal * 255).astype(np.uint8) if sal.max() <= 1.0 else sal.astype(np.uint8)
    otsu, _ = cv2.threshold(u8, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    otsu = otsu / 255.0
    hist, edges = np.histogram(sal.flatten(), 50, range=(0, 1))
    share = np.cumsum(hist)
    share = share / share[-1]
    by_share = edges[np.argmax(share > 0.7)]
    return max(0.05, min(0.5, min(otsu, by_share)))


def create_saliency_mask(saliency_map, threshold=None, smooth=True):
    """GAN_functions.py:159-208.  smooth=True: bilateral(9,75,75) -> Gaussian 31x31 -> / max (the threshold is computed and not
    used in that mode, App. D.5 - skipped here, it has no effect on the result).  smooth=False: binary mask at `threshold`, or
    at the adaptive threshold of :172-194 when none is given."""
    import cv2
    sal = np.asarray(saliency_map)
    if smooth:
        mask = cv2.bilateralFilter(sal.astype(np.float32), 9, 75, 75)
        mask = cv2.GaussianBlur(mask, (31, 31), 0)
        peak = mask.max()
        return mask / peak if peak > 0 else mask
    final_threshold = adaptive_threshold(sal) if threshold is None else threshold
    return (sal > final_threshold).astype(np.float32)
