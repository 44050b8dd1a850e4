"""Image I/O conventions and the saliency-mask front end of the reference (GAN_functions.py:24-208).

These are the steps *before* the accelerated path: file loading and the opencv-contrib saliency
pipeline.  They follow the reference operation for operation so that masks (and therefore dt, hq_ratio, actual_bpp and the
blended image) are the reference's:

  * `compute_saliency_map` runs the two cv2.saliency detectors (opencv-contrib) and their mix on the GPU (`ops.saliency_map`);
  * `create_saliency_mask` runs on the GPU in both modes: smooth (bilateral 9/75/75, Gaussian 31x31, / max - the only mode the
    reference uses, GAN_test.py:280,553) and binary (given or adaptive Otsu / histogram-share threshold) (§8 f2).
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


def compute_saliency_map(image, method="spectral_residual"):
    """GAN_functions.py:52-121 on the GPU (`ops.saliency_map`, csrc/saliency_map.cu): the reference calls cv2.saliency
    (opencv-contrib) on the CPU; here the spectral-residual and fine-grained detectors, the RAW 0.6 / 0.4 mix of 'combined' (:95)
    and the division by the maximum (:98-99, :118-119) run on the device.  Same input conventions (:63-71), same ValueError for an
    unknown method (:110), float32 (H,W) numpy array out.  The reference's fall-backs for a detector that reports failure
    (:81-91, :112-115) have no counterpart: neither detector can fail."""
    from . import ops
    image = np.asarray(image)
    if image.ndim == 2:                                 # OpenCV's detectors take a one-channel image as is
        image = np.repeat(image[:, :, None], 3, axis=2)
    return ops.saliency_map(image, method).cpu().numpy()


def enhance_saliency_map(saliency_map):
    """GAN_functions.py:123-157 on the GPU (`ops.saliency_enhance`); the reference defines it and never calls it."""
    from . import ops
    return ops.saliency_enhance(np.asarray(saliency_map, np.float32)).cpu().numpy()


def adaptive_threshold(saliency_map):
    """The threshold create_saliency_mask derives when none is given (GAN_functions.py:172-194): min(Otsu on the uint8 map, the
    lower edge of the first of 50 histogram bins whose cumulative share exceeds 0.7), clamped to [0.05, 0.5] - on the device."""
    from . import ops
    _, thr = ops.saliency_mask_binary(np.asarray(saliency_map, np.float32), None, return_threshold=True)
    return float(thr[0].item())


def create_saliency_mask(saliency_map, threshold=None, smooth=True):
    """GAN_functions.py:159-208 on the GPU.  smooth=True: bilateral(9,75,75) -> Gaussian 31x31 -> / max (`ops.saliency_mask_smooth`;
    the threshold the reference computes first has no effect in that mode, App. D.5, and is skipped).  smooth=False: binary mask at
    `threshold`, or at the adaptive threshold of :172-194 when none is given (`ops.saliency_mask_binary`).  numpy in, numpy out."""
    from . import ops
    sal = np.asarray(saliency_map, np.float32)
    if smooth:
        return ops.saliency_mask_smooth(sal).cpu().numpy()
    return ops.saliency_mask_binary(sal, threshold).cpu().numpy()
