"""Deterministic synthetic images, saliency masks and target-bpp vectors.

The reference ships no dataset (`test_dataset/`, `dataset/` are absent) and its saliency masks
come from opencv-contrib (`GAN_functions.py:52-208`, out of scope), so every benchmark and
parity test runs on synthetic inputs.  Images are produced by integer-only arithmetic (a
counter hash of (seed, image, y, x, c) plus low-frequency integer terms) so that the bytes do
not depend on any floating-point library.  Masks mimic `create_saliency_mask(smooth=True)`
(`GAN_functions.py:199-203`): smooth blobs in [0, 1] whose maximum is exactly 1.

Pixel conventions:
  * autoencoder: BGR uint8 / 255           (`test_autoencoder.py:22-25`)
  * GAN codec:   RGB (uint8 - 127.5)/127.5 (`GAN_functions.py:31-37`)
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 42  # the reference's seed (`GAN_train.py:25-26`)

_M32 = np.uint64(0xFFFFFFFF)


def _mix32(h: np.ndarray) -> np.ndarray:
    """murmur3 32-bit finaliser on uint64 arrays holding 32-bit values."""
    h = h & _M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    h ^= h >> np.uint64(16)
    return h


def synth_images_u8(n: int, height: int, width: int, seed: int = SEED_BASE, first_index: int = 0,
                    channels: int = 3) -> np.ndarray:
    """(n, H, W, C) uint8 images: slanted integer gradients + blocks + hash noise (|noise| <= 16).

    `first_index` lets each rank of a sharded job generate its own slice of the global batch
    (image i of rank r is global image first_index + i) with no communication.
    """
    idx = (np.arange(n, dtype=np.int64) + first_index)[:, None, None, None]
    y = np.arange(height, dtype=np.int64)[None, :, None, None]
    x = np.arange(width, dtype=np.int64)[None, None, :, None]
    c = np.arange(channels, dtype=np.int64)[None, None, None, :]

    # low-frequency term: triangle wave of a slanted coordinate, period 512 px
    t = (3 * x + 2 * y + 37 * idx + 64 * c) % 512
    tri = np.where(t < 256, t, 511 - t)                       # 0..255
    # mid-frequency term: 32-px blocks whose level depends on (block, image, channel)
    blk = ((x >> 5) * 7 + (y >> 5) * 13 + idx * 5 + c * 3) % 8  # 0..7
    # hash noise in [-16, 15]
    h = (np.uint64(seed & 0xFFFFFFFF)
         ^ (idx.astype(np.uint64) * np.uint64(0x9E3779B1) & _M32)
         ^ (y.astype(np.uint64) * np.uint64(0x85EBCA77) & _M32)
         ^ (x.astype(np.uint64) * np.uint64(0xC2B2AE3D) & _M32)
         ^ (c.astype(np.uint64) * np.uint64(0x27D4EB2F) & _M32))
    noise = (_mix32(h) >> np.uint64(27)).astype(np.int64) - 16
    v = 128 + ((tri - 128) * 3) // 4 + (blk - 4) * 6 + noise
    return np.clip(v, 0, 255).astype(np.uint8)


def to_unit_range(img_u8: np.ndarray) -> np.ndarray:
    """Autoencoder convention: float32 in [0, 1] (`test_autoencoder.py:24`)."""
    return img_u8.astype(np.float32) / np.float32(255.0)


def to_signed_range(img_u8: np.ndarray) -> np.ndarray:
    """GAN convention: float32 in [-1, 1] (`GAN_functions.py:37`)."""
    return (img_u8.astype(np.float32) - np.float32(127.5)) / np.float32(127.5)


def synth_masks(n: int, height: int, width: int, seed: int = SEED_BASE, first_index: int = 0) -> np.ndarray:
    """(n, H, W, 1) float32 saliency masks in [0, 1] with max exactly 1 per image.

    Each mask is a sum of 1-3 isotropic Gaussian blobs whose centres, widths and amplitudes are
    integer functions of (seed, image); blob widths scale with the image size so that the
    HQ-region ratio falls in the band published in `hq_ratio_by_bpp.png` (0.02-0.2 for target
    bpp 0.1-2.0).
    """
    out = np.empty((n, height, width, 1), dtype=np.float32)
    yy = np.arange(height, dtype=np.float64)[:, None]
    xx = np.arange(width, dtype=np.float64)[None, :]
    short = min(height, width)
    for i in range(n):
        g = first_index + i
        h = int(_mix32(np.array([(seed * 0x9E3779B1 + g * 0x85EBCA77 + 0x165667B1) & 0xFFFFFFFF], dtype=np.uint64))[0])
        nblob = 1 + (h % 3)
        acc = np.zeros((height, width), dtype=np.float64)
        for b in range(nblob):
            hb = int(_mix32(np.array([(h + 0x27D4EB2F * (b + 1)) & 0xFFFFFFFF], dtype=np.uint64))[0])
            cy = (hb & 0xFF) * height // 256
            cx = ((hb >> 8) & 0xFF) * width // 256
            sig = short * (20 + ((hb >> 16) & 0x1F)) // 256        # 0.078..0.2 of the short side
            sig = max(sig, 2)
            amp = (160 + ((hb >> 21) & 0x5F)) / 255.0                # 0.63..1.0
            acc += amp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2.0 * sig * sig))
        acc /= acc.max()
        out[i, :, :, 0] = acc.astype(np.float32)
        # guarantee max == 1 exactly after the float32 cast
        out[i, :, :, 0].flat[np.argmax(out[i, :, :, 0])] = np.float32(1.0)
    return out


def rate_control_bpps() -> np.ndarray:
    """The ten sweep targets of `GAN_test.py:534` (np.linspace(0.1, 2.0, 10))."""
    return np.linspace(0.1, 2.0, 10)
