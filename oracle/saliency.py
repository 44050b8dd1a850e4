"""CPU restatement of the two opencv-contrib saliency detectors behind `compute_saliency_map`
(GAN_functions.py:52-121).  TEST INFRASTRUCTURE.

The reference calls `cv2.saliency.StaticSaliencySpectralResidual_create().computeSaliency(bgr_u8)` and
`cv2.saliency.StaticSaliencyFineGrained_create().computeSaliency(bgr_u8)` (GAN_functions.py:75-79,102-108).  `cv2.saliency`
belongs to opencv-contrib, which is NOT installed here (plain OpenCV 4.13), so both detectors are restated from their published
source (opencv_contrib 4.x, modules/saliency/src/staticSaliencySpectralResidual.cpp and staticSaliencyFineGrained.cpp).
PARITY UNPINNED for the composition; what IS pinned:

  * every OpenCV *core* routine those two files call is importable, so each detector exists here twice: `*_cv` composes the REAL
    core routines (`cv2.cvtColor`, `cv2.resize(INTER_LINEAR_EXACT / INTER_LINEAR)`, `cv2.dft`, `cv2.cartToPolar`, `cv2.log`,
    `cv2.blur`, `cv2.exp`, `cv2.polarToCart`, `cv2.GaussianBlur`, `cv2.integral`) in the order of the contrib source, and `*_np`
    is the plain-numpy restatement the CUDA kernels follow (own bit-exact resize, box / Gaussian filters, DFT, integral image);
  * tests/test_oracle_saliency.py checks the numpy pieces against the real routines: the integer ones bit for bit, the
    floating-point ones to 1e-6 / 1e-3 (cv2's cartToPolar / polarToCart work in float32 with a polynomial arctangent that is
    1.6e-4 rad off; the restatement and the kernels use exact double arithmetic there, which is the larger part of the 1e-3).

Fine-grained (Montabone & Soto) returns, in opencv-contrib 4.x, the uint8 intensity-conspicuity map converted to float32 / 255.
"""
from __future__ import annotations

import numpy as np

from .metrics import bgr2gray_u8

SR_SIZE = 64                                            # resImWidth = resImHeight = 64 (StaticSaliencySpectralResidual ctor)
FG_NEIGHBORHOODS = (12, 24, 48, 28, 56, 112)            # 3*4, 3*4*2, 3*4*2*2, 7*4, 7*4*2, 7*4*2*2 (calcIntensityChannel)


# ------------------------------------------------------------------ shared pieces (numpy restatements of OpenCV core routines)

def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    """BORDER_REFLECT_101 (BORDER_DEFAULT): gfedcb|abcdefgh|gfedcba."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def _exact_coeffs(src: int, dst: int):
    """interpolationLinear<ufixedpoint16>::getCoeffs (imgproc/src/resize.cpp, resize_bitExact): 8 fractional bits, the first /
    last destination pixels whose source position falls outside take the border pixel with weight 1."""
    scale = src / dst                                   # softdouble::one() / softdouble(inv_scale): both are exact IEEE doubles
    ofs = np.zeros(dst, np.int64)
    c0 = np.full(dst, 256, np.int64)
    c1 = np.zeros(dst, np.int64)
    for d in range(dst):
        f = scale * (d + 0.5) - 0.5
        i = int(np.floor(f))
        if i >= 0 and src > 1:
            if i < src - 1:
                ofs[d] = i
                c1[d] = int(np.rint((f - i) * 256))     # cvRound = round half to even
                c0[d] = 256 - c1[d]
            else:
                ofs[d] = src - 1                        # weight 1 on the last pixel
        # i < 0: weight 1 on pixel 0 (ofs 0)
    return ofs, c0, c1


def resize_linear_exact_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR_EXACT) for one-channel uint8: horizontal pass to 8.8 fixed
    point, vertical pass to 16.16, round half up."""
    h, w = img.shape
    ox, cx0, cx1 = _exact_coeffs(w, dst_w)
    oy, cy0, cy1 = _exact_coeffs(h, dst_h)
    s = img.astype(np.int64)
    ox1 = np.minimum(ox + 1, w - 1)
    rows = s[:, ox] * cx0 + s[:, ox1] * cx1             # (h, dst_w), <= 255 * 256
    oy1 = np.minimum(oy + 1, h - 1)
    out = rows[oy] * cy0[:, None] + rows[oy1] * cy1[:, None]
    return ((out + (1 << 15)) >> 16).astype(np.uint8)


def resize_linear_f32(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for one-channel float32 (imgproc/src/resize.cpp,
    resizeGeneric_ + HResizeLinear + VResizeLinear): float32 coefficients, rows first."""
    h, w = img.shape

    def coeffs(src, dst):
        scale = 1.0 / (dst / src)
        f = ((np.arange(dst) + 0.5) * scale - 0.5).astype(np.float32)
        i = np.floor(f).astype(np.int64)
        f = f - i.astype(np.float32)
        lo = i < 0
        f[lo], i[lo] = 0, 0
        hi = i >= src - 1
        f[hi], i[hi] = 0, src - 1
        return i, (np.float32(1) - f).astype(np.float32), f.astype(np.float32)

    ix, ax0, ax1 = coeffs(w, dst_w)
    iy, ay0, ay1 = coeffs(h, dst_h)
    s = img.astype(np.float32)
    rows = s[:, ix] * ax0 + s[:, np.minimum(ix + 1, w - 1)] * ax1
    return (rows[iy] * ay0[:, None] + rows[np.minimum(iy + 1, h - 1)] * ay1[:, None]).astype(np.float32)


def gaussian3_u8(img: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(u8, (3, 3), 0): the fixed-point path, kernel [1 2 1] / 4 both ways, BORDER_REFLECT_101, round half up."""
    h, w = img.shape
    s = img.astype(np.int64)
    xs = np.arange(w)
    hsum = s[:, _reflect101(xs - 1, w)] + 2 * s + s[:, _reflect101(xs + 1, w)]
    ys = np.arange(h)
    v = hsum[_reflect101(ys - 1, h)] + 2 * hsum + hsum[_reflect101(ys + 1, h)]
    return ((v + 8) >> 4).astype(np.uint8)


def integral_f32(img_u8: np.ndarray) -> np.ndarray:
    """cv2.integral(u8, sdepth=CV_32F): (h+1, w+1) float32, sum[y+1][x+1] = sum[y][x+1] + (running float32 sum of row y up to x)
    (imgproc/src/sumpixels.cpp, integral_).  Row sums of uint8 stay exact integers in float32 for w <= 65793."""
    h, w = img_u8.shape
    out = np.zeros((h + 1, w + 1), np.float32)
    row = np.cumsum(img_u8.astype(np.float64), axis=1).astype(np.float32)
    for y in range(h):
        out[y + 1, 1:] = out[y, 1:] + row[y]            # float32 addition, one rounding per element
    return out


def _box3_reflect101(x: np.ndarray) -> np.ndarray:
    """cv2.blur(x, (3, 3)) on float64: normalised box filter, BORDER_REFLECT_101."""
    h, w = x.shape
    xs, ys = np.arange(w), np.arange(h)
    hs = x[:, _reflect101(xs - 1, w)] + x + x[:, _reflect101(xs + 1, w)]
    return (hs[_reflect101(ys - 1, h)] + hs + hs[_reflect101(ys + 1, h)]) * (1.0 / 9.0)


def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma) for sigma > 0 (float64)."""
    x = np.arange(ksize) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return k / k.sum()


def _gauss5_reflect101(x: np.ndarray, sigma: float) -> np.ndarray:
    k = gaussian_kernel(5, sigma)
    h, w = x.shape
    xs, ys = np.arange(w), np.arange(h)
    hs = sum(k[t] * x[:, _reflect101(xs + t - 2, w)] for t in range(5))
    return sum(k[t] * hs[_reflect101(ys + t - 2, h)] for t in range(5))


# ------------------------------------------------------------------ spectral residual (Hou & Zhang 2007)

def spectral_residual_np(bgr_u8: np.ndarray) -> np.ndarray:
    """StaticSaliencySpectralResidual::computeSaliencyImpl: gray -> 64x64 (LINEAR_EXACT) -> DFT -> log amplitude minus its 3x3
    mean -> back with the original phase -> |.| -> Gaussian 5x5 sigma 8 -> square -> / max -> float32 -> resize to the image."""
    h, w = bgr_u8.shape[:2]
    gray = bgr2gray_u8(bgr_u8) if bgr_u8.ndim == 3 else bgr_u8
    small = resize_linear_exact_u8(gray, SR_SIZE, SR_SIZE).astype(np.float64)
    spec = np.fft.fft2(small)
    amp = np.abs(spec)
    with np.errstate(divide="ignore", invalid="ignore"):
        log_amp = np.log(amp)
        gain = np.exp(log_amp - _box3_reflect101(log_amp))          # new amplitude; new spectrum = gain * e^{i phase}
        back = np.fft.ifft2(np.where(amp > 0, spec / amp, 1.0) * gain) * (SR_SIZE * SR_SIZE)   # cv2.dft(DFT_INVERSE) does not scale
    mag = _gauss5_reflect101(np.abs(back), 8.0)
    mag = mag * mag
    mag = mag / mag.max()
    return resize_linear_f32(mag.astype(np.float32), h, w)


def spectral_residual_cv(bgr_u8: np.ndarray) -> np.ndarray:
    """The same function composed of the real OpenCV core calls, in the order of the contrib source."""
    import cv2
    h, w = bgr_u8.shape[:2]
    gray = cv2.cvtColor(bgr_u8, cv2.COLOR_BGR2GRAY) if bgr_u8.ndim == 3 else bgr_u8
    small = cv2.resize(gray, (SR_SIZE, SR_SIZE), interpolation=cv2.INTER_LINEAR_EXACT)
    real = small.astype(np.float64)
    spec = cv2.dft(cv2.merge([real, np.zeros_like(real)]))
    re, im = cv2.split(spec)
    magnitude, angle = cv2.cartToPolar(re, im, angleInDegrees=False)
    log_amp = cv2.log(magnitude)
    log_blur = cv2.blur(log_amp, (3, 3), borderType=cv2.BORDER_DEFAULT)
    magnitude = cv2.exp(log_amp - log_blur)
    re, im = cv2.polarToCart(magnitude, angle, angleInDegrees=False)
    back = cv2.dft(cv2.merge([re, im]), flags=cv2.DFT_INVERSE)
    re, im = cv2.split(back)
    magnitude, _ = cv2.cartToPolar(re, im, angleInDegrees=False)
    magnitude = cv2.GaussianBlur(magnitude, (5, 5), 8, None, 0, cv2.BORDER_DEFAULT)
    magnitude = magnitude * magnitude
    magnitude = magnitude / magnitude.max()
    return cv2.resize(magnitude.astype(np.float32), (w, h), interpolation=cv2.INTER_LINEAR)


# ------------------------------------------------------------------ fine grained (Montabone & Soto 2010)

def _fg_from_blurred(gray: np.ndarray, integral: np.ndarray) -> np.ndarray:
    """getIntensityScaled / getMean / mixScales / mixOnOff on the twice-blurred gray image and its float32 integral image."""
    h, w = gray.shape
    ys, xs = np.mgrid[0:h, 0:w]
    g = gray.astype(np.float32)
    on_sum = np.zeros((h, w), np.int64)
    off_sum = np.zeros((h, w), np.int64)
    for nb in FG_NEIGHBORHOODS:
        # getMean: window corners in integral-image coordinates, clamped to [0, cols-1] x [0, rows-1] of the (h+1, w+1) image
        x1 = np.clip(xs - nb + 1, 0, w)
        y1 = np.clip(ys - nb + 1, 0, h)
        x2 = np.clip(xs + nb + 1, 0, w)
        y2 = np.clip(ys + nb + 1, 0, h)
        v = ((integral[y2, x2] + integral[y1, x1]) - integral[y2, x1]) - integral[y1, x2]      # float32, that order
        area = ((x2 - x1) * (y2 - y1) - 1).astype(np.float32)
        mean = (v - g) / area                                                                  # float32
        on = g - mean
        off = mean - g
        on_sum += np.where(on > 0, on, 0).astype(np.uint8)          # (uchar) of a float in [0, 255]: truncation
        off_sum += np.where(off > 0, off, 0).astype(np.uint8)

    def norm(total):                                    # mixScales: (uchar)(255. * (float)(sum / (float)max_sum))
        peak = int(total.max())
        if peak == 0:
            return np.zeros((h, w), np.uint8)
        q = total.astype(np.float32) / np.float32(peak)
        return (255.0 * q.astype(np.float64)).astype(np.uint8)

    on_u8, off_u8 = norm(on_sum), norm(off_sum)
    peak = max(int(on_u8.max()), int(off_u8.max()))     # mixOnOff
    if peak == 0:
        return np.zeros((h, w), np.uint8)
    both = (on_u8.astype(np.int32) + off_u8).astype(np.float32)
    mix = 255.0 * both.astype(np.float64) / np.float64(np.float32(peak))       # 255. * (float)(on + off) / (float)maxVal: double
    # (uchar) of a double that can exceed 255 (a pixel may be "on" at one scale and "off" at another): x86 converts to int32 and
    # keeps the low byte
    return (mix.astype(np.int32) & 0xFF).astype(np.uint8)


def fine_grained_np(bgr_u8: np.ndarray) -> np.ndarray:
    """StaticSaliencyFineGrained::computeSaliencyImpl -> float32 map in [0, 1] (uint8 conspicuity / 255)."""
    gray = bgr2gray_u8(bgr_u8) if bgr_u8.ndim == 3 else bgr_u8
    gray = gaussian3_u8(gaussian3_u8(gray))
    return _fg_from_blurred(gray, integral_f32(gray)).astype(np.float32) * np.float32(1.0 / 255.0)


def fine_grained_cv(bgr_u8: np.ndarray) -> np.ndarray:
    """The same with cv2.cvtColor / cv2.GaussianBlur / cv2.integral doing the library's part."""
    import cv2
    gray = cv2.cvtColor(bgr_u8, cv2.COLOR_BGR2GRAY) if bgr_u8.ndim == 3 else bgr_u8.copy()
    gray = cv2.GaussianBlur(gray, (3, 3), 0, None, 0)
    gray = cv2.GaussianBlur(gray, (3, 3), 0, None, 0)
    integral = cv2.integral(gray, sdepth=cv2.CV_32F)
    return _fg_from_blurred(gray, integral).astype(np.float32) * np.float32(1.0 / 255.0)


# ------------------------------------------------------------------ compute_saliency_map (GAN_functions.py:52-121)

def to_cv_bgr_u8(image: np.ndarray) -> np.ndarray:
    """GAN_functions.py:63-71."""
    image = np.asarray(image)
    if image.dtype == np.float32 and np.max(image) <= 1.0:
        image_cv = ((image + 1) * 127.5).astype(np.uint8)
    else:
        image_cv = image.astype(np.uint8)
    if image_cv.ndim == 3 and image_cv.shape[2] == 3:
        image_cv = image_cv[:, :, ::-1]
    return np.ascontiguousarray(image_cv)


def compute_saliency_map(image: np.ndarray, method: str = "spectral_residual", use_cv: bool = False) -> np.ndarray:
    """GAN_functions.py:52-121 with the detectors above (both always succeed)."""
    bgr = to_cv_bgr_u8(image)
    sr = spectral_residual_cv if use_cv else spectral_residual_np
    fg = fine_grained_cv if use_cv else fine_grained_np
    if method == "combined":
        out = 0.6 * sr(bgr) + 0.4 * fg(bgr)             # float64 = python float * float32 array under numpy 2 -> stays float32
    elif method == "spectral_residual":
        out = sr(bgr)
    elif method == "fine_grained":
        out = fg(bgr)
    else:
        raise ValueError(f"Unsupported saliency method: {method}")
    peak = out.max()
    return out / peak if peak > 0 else out


# ------------------------------------------------------------------ create_saliency_mask (GAN_functions.py:159-208)

def otsu_u8(u8: np.ndarray) -> int:
    """getThreshVal_Otsu_8u (imgproc/src/thresh.cpp): the threshold cv2.threshold(u8, 0, 255, THRESH_BINARY + THRESH_OTSU) returns.
    Pinned on the real call in tests/test_oracle_saliency.py."""
    h = np.bincount(u8.ravel(), minlength=256).astype(np.int64)
    scale = 1.0 / u8.size
    mu = 0.0
    for i in range(256):
        mu += i * float(h[i])
    mu *= scale
    mu1 = q1 = max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = h[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma, max_val = sigma, i
    return max_val


def adaptive_threshold(saliency_map: np.ndarray, use_cv: bool = True) -> float:
    """GAN_functions.py:172-194: min(Otsu on the uint8 map / 255, lower edge of the first of 50 histogram bins whose cumulative
    share exceeds 0.7), clamped to [0.05, 0.5]."""
    sal = np.asarray(saliency_map)
    u8 = (sal * 255).astype(np.uint8) if sal.max() <= 1.0 else sal.astype(np.uint8)
    if use_cv:
        import cv2
        otsu, _ = cv2.threshold(u8, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    else:
        otsu = otsu_u8(u8)
    otsu = otsu / 255.0
    hist, edges = np.histogram(sal.flatten(), 50, range=(0, 1))
    share = np.cumsum(hist)
    share = share / share[-1]
    by_share = edges[np.argmax(share > 0.7)]
    return float(max(0.05, min(0.5, min(otsu, by_share))))


def create_saliency_mask(saliency_map: np.ndarray, threshold=None, smooth: bool = True) -> np.ndarray:
    """GAN_functions.py:159-208 with the real OpenCV calls (the threshold is computed and unused when smooth=True, App. D.5)."""
    import cv2
    sal = np.asarray(saliency_map)
    if smooth:
        mask = cv2.bilateralFilter(sal.astype(np.float32), 9, 75, 75)
        mask = cv2.GaussianBlur(mask, (31, 31), 0)
        peak = mask.max()
        return mask / peak if peak > 0 else mask
    final_threshold = adaptive_threshold(sal) if threshold is None else threshold
    return (sal > final_threshold).astype(np.float32)


def enhance_saliency_map(saliency_map: np.ndarray) -> np.ndarray:
    """GAN_functions.py:123-157 with the real OpenCV calls."""
    import cv2
    filtered = cv2.bilateralFilter(saliency_map.astype(np.float32), 9, 75, 75)
    enhanced = np.zeros_like(saliency_map)
    for scale, weight in zip((3, 9, 15), (0.5, 0.3, 0.2)):
        enhanced += weight * cv2.GaussianBlur(filtered, (scale, scale), 0)
    return np.clip(np.power(enhanced, 0.8), 0, 1)
