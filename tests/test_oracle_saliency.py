"""oracle/saliency.py: the numpy restatements of the OpenCV core routines inside the two cv2.saliency detectors, checked against the
REAL routines (cv2 core is installed; the contrib module that composes them is not - the composition is restated from its source)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import saliency as S  # noqa: E402

SIZES = [(256, 256), (100, 173), (64, 64), (300, 40), (65, 63), (32, 48), (540, 960)]


def photo(h, w, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 90 * np.sin(x / 23.0 + k) * np.cos(y / 31.0 - k) for k in range(3)], -1)
    blob = 80 * np.exp(-((x - w * 0.6) ** 2 + (y - h * 0.4) ** 2) / (2 * (min(h, w) / 8) ** 2))[..., None]
    return np.clip(base + blob + rng.integers(-20, 20, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w", SIZES)
def test_integer_pieces_are_bit_exact(h, w):
    g = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w), dtype=np.uint8)
    np.testing.assert_array_equal(S.resize_linear_exact_u8(g, 64, 64), cv2.resize(g, (64, 64), interpolation=cv2.INTER_LINEAR_EXACT))
    np.testing.assert_array_equal(S.gaussian3_u8(g), cv2.GaussianBlur(g, (3, 3), 0))
    np.testing.assert_array_equal(S.integral_f32(g), cv2.integral(g, sdepth=cv2.CV_32F))


@pytest.mark.parametrize("h,w", SIZES)
def test_float_pieces(h, w):
    rng = np.random.default_rng(3)
    m = rng.random((64, 64), dtype=np.float32)
    np.testing.assert_allclose(S.resize_linear_f32(m, h, w), cv2.resize(m, (w, h), interpolation=cv2.INTER_LINEAR), atol=5e-6)
    d = rng.standard_normal((64, 64))
    np.testing.assert_allclose(S._box3_reflect101(d), cv2.blur(d, (3, 3), borderType=cv2.BORDER_DEFAULT), atol=1e-12)
    np.testing.assert_allclose(S._gauss5_reflect101(d, 8.0), cv2.GaussianBlur(d, (5, 5), 8, None, 0, cv2.BORDER_DEFAULT), atol=1e-12)
    np.testing.assert_allclose(S.gaussian_kernel(5, 8.0), cv2.getGaussianKernel(5, 8.0, cv2.CV_64F).ravel(), atol=1e-15)
    spec = cv2.dft(cv2.merge([d, np.zeros_like(d)]))
    np.testing.assert_allclose(spec[..., 0] + 1j * spec[..., 1], np.fft.fft2(d), atol=1e-9)


@pytest.mark.parametrize("h,w", [(256, 256), (100, 173), (512, 384)])
def test_detectors_numpy_vs_real_core_routines(h, w):
    bgr = photo(h, w, seed=h)
    np.testing.assert_allclose(S.spectral_residual_np(bgr), S.spectral_residual_cv(bgr), atol=1e-3)   # float32 polar round trip in cv2
    np.testing.assert_array_equal(S.fine_grained_np(bgr), S.fine_grained_cv(bgr))
    rgb = (bgr[:, :, ::-1].astype(np.float32) - 127.5) / 127.5
    for method in ("spectral_residual", "fine_grained", "combined"):
        a = S.compute_saliency_map(rgb, method)
        assert a.dtype == np.float32 and a.shape == (h, w) and a.max() == pytest.approx(1.0) and a.min() >= 0
        np.testing.assert_allclose(a, S.compute_saliency_map(rgb, method, use_cv=True), atol=1e-3)
    with pytest.raises(ValueError, match="Unsupported"):
        S.compute_saliency_map(rgb, "nope")


def test_fine_grained_known_properties():
    """A flat image has no centre-surround contrast anywhere: both polarity maps are zero and the detector returns zeros (no
    division by the zero maximum); a bright square on a dark ground is 'on' inside and 'off' in a halo around it (each polarity is normalised by its own maximum), fading with distance."""
    flat = np.full((64, 64, 3), 77, np.uint8)
    assert S.fine_grained_np(flat).max() == 0
    img = np.full((96, 96, 3), 20, np.uint8)
    img[40:56, 40:56] = 220
    fg = S.fine_grained_np(img)
    assert fg[48, 48] > 0.9 and fg[36, 36] > 0.7 and fg[5, 5] < 0.4 and fg.max() == pytest.approx(1.0)


@pytest.mark.skipif(not hasattr(cv2, "saliency"), reason="opencv-contrib (cv2.saliency) is not installed: the composition stays unpinned")
@pytest.mark.parametrize("h,w", [(256, 256), (100, 173)])
def test_against_opencv_contrib_when_present(h, w):
    bgr = photo(h, w, seed=1)
    ok, sr = cv2.saliency.StaticSaliencySpectralResidual_create().computeSaliency(bgr)
    assert ok
    np.testing.assert_allclose(S.spectral_residual_np(bgr), sr, atol=1e-3)
    ok, fg = cv2.saliency.StaticSaliencyFineGrained_create().computeSaliency(bgr)
    assert ok
    np.testing.assert_allclose(S.fine_grained_np(bgr), fg, atol=1.5 / 255)


def test_otsu_restatement_matches_opencv():
    """oracle.saliency.otsu_u8 (getThreshVal_Otsu_8u restated; the CUDA threshold kernel follows it) against cv2.threshold."""
    rng = np.random.default_rng(0)
    for t in range(120):
        h, w = rng.integers(8, 160, 2)
        kind = t % 4
        if kind == 0:
            m = rng.random((h, w)) ** rng.uniform(0.3, 4)
        elif kind == 1:
            m = (rng.random((h, w)) > rng.random()) * rng.random()
        elif kind == 2:
            m = np.clip(rng.normal(rng.random(), 0.2, (h, w)), 0, 1)
        else:
            m = np.round(rng.random((h, w)) * rng.integers(2, 9)) / 8
        u8 = (m.astype(np.float32) * 255).astype(np.uint8)
        want, _ = cv2.threshold(u8, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert S.otsu_u8(u8) == int(want)
        assert S.adaptive_threshold(m.astype(np.float32), use_cv=False) == S.adaptive_threshold(m.astype(np.float32), use_cv=True)
