"""numpy/scipy restatement of the reference's evaluation arithmetic.  TEST INFRASTRUCTURE.

Follows `GAN_functions.py:724-759` (compute_metrics), `test_autoencoder.py:49-66`
(calculate_mse/psnr/ssim) and `GAN_test.py:310-325,573-582` (bpp accounting).  scikit-image is
not installed, so `peak_signal_noise_ratio` / `structural_similarity` are restated from their
published algorithm (skimage >= 0.19); `scipy.ndimage.uniform_filter`, the
routine scikit-image itself calls, is used directly.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def _float_type(*arrays):
    """skimage._shared.utils._supported_float_type: float32/float16 -> float32, else float64."""
    if all(a.dtype in (np.float32, np.float16) for a in arrays):
        return np.float32
    return np.float64


def sk_psnr(image_true: np.ndarray, image_test: np.ndarray, data_range: float) -> float:
    ft = _float_type(image_true, image_test)
    a = image_true.astype(ft, copy=False)
    b = image_test.astype(ft, copy=False)
    err = np.mean((a - b) ** 2, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return float(10 * np.log10((data_range ** 2) / err))


def sk_ssim(im1: np.ndarray, im2: np.ndarray, data_range: float, win_size: int = 7,
            k1: float = 0.01, k2: float = 0.03, full: bool = False):
    """structural_similarity defaults: uniform 7x7 window, sample covariance, crop (win-1)//2."""
    assert im1.ndim == 2 and im1.shape == im2.shape
    ft = _float_type(im1, im2)
    a = im1.astype(ft, copy=False)
    b = im2.astype(ft, copy=False)
    npix = win_size ** a.ndim
    cov_norm = npix / (npix - 1)
    ux = uniform_filter(a, size=win_size)
    uy = uniform_filter(b, size=win_size)
    uxx = uniform_filter(a * a, size=win_size)
    uyy = uniform_filter(b * b, size=win_size)
    uxy = uniform_filter(a * b, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (k1 * data_range) ** 2
    c2 = (k2 * data_range) ** 2
    a1, a2, b1, b2 = (2 * ux * uy + c1, 2 * vxy + c2, ux ** 2 + uy ** 2 + c1, vx + vy + c2)
    s = (a1 * a2) / (b1 * b2)
    pad = (win_size - 1) // 2
    mssim = float(s[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return (mssim, s) if full else mssim


def compute_metrics(original_img: np.ndarray, compressed_img: np.ndarray) -> dict:
    """GAN_functions.py:724-759 - inputs in [-1,1], float32 (H,W,3)."""
    o = (original_img + 1) / 2
    c = (compressed_img + 1) / 2
    psnr_value = sk_psnr(o, c, 1.0)                                                    # :740
    if original_img.ndim == 3 and original_img.shape[2] == 3:
        ssim_value = float(np.mean([sk_ssim(o[:, :, i], c[:, :, i], 1.0) for i in range(3)]))  # :745-748
    else:
        ssim_value = sk_ssim(o, c, 1.0)
    mse_value = np.mean((o - c) ** 2)                                                  # :753
    return {"psnr": psnr_value, "ssim": ssim_value, "mse": mse_value}


def bgr2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(BGR2GRAY) on uint8 in OpenCV 4.x: 15-bit fixed point, rounded."""
    b = img[..., 0].astype(np.int64)
    g = img[..., 1].astype(np.int64)
    r = img[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def ae_calculate_mse(image1: np.ndarray, image2: np.ndarray) -> float:
    """test_autoencoder.py:49-50 - on uint8 operands the difference and the square wrap mod 256."""
    return float(np.mean((image1 - image2) ** 2))


def ae_true_mse(image1: np.ndarray, image2: np.ndarray) -> float:
    d = image1.astype(np.float64) - image2.astype(np.float64)
    return float(np.mean(d * d))


def ae_calculate_psnr(image1: np.ndarray, image2: np.ndarray) -> float:
    """test_autoencoder.py:52-57."""
    dr = 255 if (image1.dtype == np.uint8 or image2.dtype == np.uint8) else 1.0
    return sk_psnr(image1, image2, dr)


def ae_calculate_ssim(image1: np.ndarray, image2: np.ndarray) -> float:
    """test_autoencoder.py:59-66 - SSIM of the BGR2GRAY images (uint8 -> float64 in skimage)."""
    if image1.dtype == np.uint8 or image2.dtype == np.uint8:
        return sk_ssim(bgr2gray_u8(image1), bgr2gray_u8(image2), 255)
    # float path: cv2 float gray = 0.114 B + 0.587 G + 0.299 R
    coef = np.array([0.114, 0.587, 0.299], dtype=np.float32)
    return sk_ssim((image1 * coef).sum(-1).astype(np.float32), (image2 * coef).sum(-1).astype(np.float32), 1.0)


def bpp_accounting(bit_alloc_map: np.ndarray, img_size=(256, 256), base_latent_dim=512) -> dict:
    """GAN_test.py:310-325 (same arithmetic at :573-582)."""
    hq_ratio = np.mean(bit_alloc_map)
    lq_ratio = 1.0 - hq_ratio
    hq_bits = hq_ratio * (base_latent_dim * 2) * 32
    lq_bits = lq_ratio * base_latent_dim * 32
    total_bits = hq_bits + lq_bits
    original_bits = img_size[0] * img_size[1] * 3 * 8
    return {"hq_ratio": hq_ratio, "lq_ratio": lq_ratio, "total_bits": total_bits,
            "compression_ratio": original_bits / total_bits,
            "actual_bpp": total_bits / (img_size[0] * img_size[1])}


def symbol_entropy_bits(symbols: np.ndarray) -> float:
    """Zeroth-order empirical entropy (bits) of an integer symbol vector.  No reference
    counterpart (the reference's bpp is analytic); parity unpinned, numpy restatement only."""
    _, counts = np.unique(np.asarray(symbols).astype(np.int64), return_counts=True)
    p = counts / counts.sum()
    return float(-(p * np.log2(p)).sum() * counts.sum())


# ---- MS-SSIM (BASELINE configs[4] "PSNR/MS-SSIM per image") -----------------------------------------------------------------
# The reference computes single-scale SSIM only (GAN_functions.py:745-748); MS-SSIM has no call site there, so this is the
# published algorithm (Wang, Simoncelli, Bovik 2003) in the form the common implementations use (tf.image.ssim_multiscale,
# pytorch_msssim): five scales, weights below, 11x11 Gaussian window sigma 1.5 applied as a VALID separable correlation, K1 =
# 0.01, K2 = 0.03, population (co)variances, 2x2 average pooling between scales, per channel
#   msssim_c = prod_{j<5} relu(mean cs_j)^w_j * relu(mean ssim_5)^w_5,   result = mean over channels.
# PARITY UNPINNED: float64 numpy restatement of that definition; the CUDA kernel is compared against it only.
MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def gaussian_window(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    x = np.arange(size, dtype=np.float64) - size // 2
    g = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return g / g.sum()


def _valid_gauss(x: np.ndarray, g: np.ndarray) -> np.ndarray:
    from scipy.ndimage import correlate1d
    r = len(g) // 2
    y = correlate1d(correlate1d(x, g, axis=0, mode="constant"), g, axis=1, mode="constant")
    return y[r:-r, r:-r]


def ms_ssim(img1: np.ndarray, img2: np.ndarray, data_range: float = 1.0) -> float:
    """MS-SSIM of two (H,W,C) images in [0, data_range]; min(H, W) > 160."""
    a, b = np.asarray(img1, np.float64), np.asarray(img2, np.float64)
    if a.ndim == 2:
        a, b = a[..., None], b[..., None]
    g = gaussian_window()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    vals = []
    for c in range(a.shape[2]):
        x, y = a[..., c], b[..., c]
        prod = 1.0
        for j, wgt in enumerate(MSSSIM_WEIGHTS):
            mx, my = _valid_gauss(x, g), _valid_gauss(y, g)
            sxx = _valid_gauss(x * x, g) - mx * mx
            syy = _valid_gauss(y * y, g) - my * my
            sxy = _valid_gauss(x * y, g) - mx * my
            cs = (2.0 * sxy + c2) / (sxx + syy + c2)
            if j < len(MSSSIM_WEIGHTS) - 1:
                prod *= max(cs.mean(), 0.0) ** wgt
                h2, w2 = x.shape[0] // 2, x.shape[1] // 2            # 2x2 average pooling (floor)
                x = x[:2 * h2, :2 * w2].reshape(h2, 2, w2, 2).mean(axis=(1, 3))
                y = y[:2 * h2, :2 * w2].reshape(h2, 2, w2, 2).mean(axis=(1, 3))
            else:
                ss = ((2.0 * mx * my + c1) / (mx * mx + my * my + c1)) * cs
                prod *= max(ss.mean(), 0.0) ** wgt
        vals.append(prod)
    return float(np.mean(vals))
