"""Pin the metric restatement: cv2 for BGR2GRAY, brute force for the SSIM window, numpy quirks."""
import numpy as np
import pytest

from oracle import metrics


def test_bgr2gray_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    np.testing.assert_array_equal(metrics.bgr2gray_u8(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    edge = np.array([[[255, 255, 255], [0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255]]], np.uint8)
    np.testing.assert_array_equal(metrics.bgr2gray_u8(edge), cv2.cvtColor(edge, cv2.COLOR_BGR2GRAY))


def test_psnr_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (40, 40, 3), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    assert abs(metrics.sk_psnr(a, b, 255) - cv2.PSNR(a, b)) < 1e-9


def _ssim_bruteforce(a, b, R):
    a = a.astype(np.float64); b = b.astype(np.float64)
    H, W = a.shape
    c1, c2 = (0.01 * R) ** 2, (0.03 * R) ** 2
    vals = []
    for y in range(3, H - 3):
        for x in range(3, W - 3):
            wa, wb = a[y - 3:y + 4, x - 3:x + 4], b[y - 3:y + 4, x - 3:x + 4]
            ux, uy = wa.mean(), wb.mean()
            vx, vy = wa.var(ddof=1), wb.var(ddof=1)
            vxy = ((wa - ux) * (wb - uy)).sum() / 48
            vals.append(((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2)))
    return float(np.mean(vals))


def test_ssim_matches_windowed_definition():
    rng = np.random.default_rng(2)
    a = rng.random((24, 31)).astype(np.float32)
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1).astype(np.float32)
    assert abs(metrics.sk_ssim(a, b, 1.0) - _ssim_bruteforce(a, b, 1.0)) < 2e-5   # float32 path
    a8 = (a * 255).astype(np.uint8); b8 = (b * 255).astype(np.uint8)
    assert abs(metrics.sk_ssim(a8, b8, 255) - _ssim_bruteforce(a8, b8, 255)) < 1e-10  # float64 path
    assert metrics.sk_ssim(a, a, 1.0) == pytest.approx(1.0, abs=1e-6)


def test_uint8_mse_wraps_like_numpy():
    a = np.array([[[10, 200, 0]]], np.uint8)
    b = np.array([[[20, 100, 255]]], np.uint8)
    # (10-20) -> 246, 246^2 = 60516 -> 100;  (200-100)=100 -> 10000 -> 16;  (0-255) -> 1 -> 1
    assert metrics.ae_calculate_mse(a, b) == pytest.approx((100 + 16 + 1) / 3)
    assert metrics.ae_true_mse(a, b) == pytest.approx((100 + 10000 + 65025) / 3)


def test_compute_metrics_shapes_and_ranges():
    rng = np.random.default_rng(3)
    a = (rng.random((32, 32, 3)).astype(np.float32) * 2 - 1)
    b = np.clip(a + 0.05 * rng.standard_normal(a.shape).astype(np.float32), -1, 1)
    m = metrics.compute_metrics(a, b)
    assert set(m) == {"psnr", "ssim", "mse"}
    assert 25 < m["psnr"] < 40 and 0 < m["ssim"] < 1 and m["mse"].dtype == np.float32


def test_symbol_entropy():
    assert metrics.symbol_entropy_bits(np.zeros(64)) == 0.0
    assert metrics.symbol_entropy_bits(np.array([0, 1] * 32)) == pytest.approx(64.0)
    assert metrics.symbol_entropy_bits(np.arange(8)) == pytest.approx(24.0)
