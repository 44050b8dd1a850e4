"""GPU parity at the BASELINE.json image sizes (configs[2..4]): 1024x1024 ROI sweep through the tiled GAN codec,
1080p and 4K frames through the autoencoder.  Where the CPU oracle finishes in seconds the comparison is direct;
at 4K a size-independent locality property is used (the network's receptive field is ~ +-16 px, so the interior of
a crop must reproduce the full-frame result)."""
import numpy as np
import pytest
import torch

from oracle import graphs, metrics
from test_gpu_models import _adaptive, _check_adaptive

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tc", "fp32"])
def precision(request, cic):
    old = cic.get_precision()
    cic.set_precision(request.param)
    yield request.param
    cic.set_precision(old)


def test_autoencoder_1080p_matches_oracle(cic, precision):
    """BASELINE configs[3]: one 1920x1080 frame (H, W divisible by 4) through build_autoencoder."""
    import train_autoencoder as tr
    H, W = 1080, 1920
    model = tr.build_autoencoder((H, W, 3))
    w = cic.weights.synthetic_autoencoder(seed=42)
    model.set_weights_dict(w)
    x = cic.synth.to_unit_range(cic.synth.synth_images_u8(1, H, W, seed=60))
    y = model.predict(x)
    want = graphs.autoencoder_forward(w, x)
    err = np.abs(y - want).max()
    assert err < (2e-5 if precision == "fp32" else 1e-2), f"max-abs {err}"
    r = cic.autoencoder.evaluate_batch(model, x)
    y8, want8 = r["compressed_u8"].cpu().numpy(), graphs.autoencoder_output_u8(want)
    assert np.abs(y8.astype(int) - want8.astype(int)).max() <= 1
    x8 = (x * 255).astype(np.uint8)
    assert abs(r["psnr"][0] - metrics.ae_calculate_psnr(x8[0], want8[0])) < 0.05          # north star: 0.05 dB
    assert abs(r["ssim"][0] - metrics.ae_calculate_ssim(x8[0], want8[0])) < 1e-3


def test_autoencoder_4k_crop_locality(cic):
    """BASELINE configs[4]: 3840x2160.  The interior of a crop (offsets multiples of 4 keep the pooling grid aligned)
    equals the same pixels of the full frame, and both agree with the oracle run on the crop."""
    import train_autoencoder as tr
    cic.set_precision("tc")
    H, W = 2160, 3840
    model = tr.build_autoencoder((H, W, 3))
    w = cic.weights.synthetic_autoencoder(seed=42)
    model.set_weights_dict(w)
    x = cic.synth.to_unit_range(cic.synth.synth_images_u8(1, H, W, seed=61))
    y = model.predict(x)
    assert y.shape == (1, H, W, 3) and y.min() > 0.0 and y.max() < 1.0
    y0, x0, ch, cw, m = 1000, 2400, 512, 768, 32
    crop = np.ascontiguousarray(x[:, y0:y0 + ch, x0:x0 + cw])
    yc = tr.build_autoencoder((ch, cw, 3))
    yc.set_weights_dict(w)
    got_c = yc.predict(crop)
    inner_full = y[:, y0 + m:y0 + ch - m, x0 + m:x0 + cw - m]
    assert np.abs(got_c[:, m:-m, m:-m] - inner_full).max() < 1e-3          # tile shapes differ: bf16 rounding flips only
    want_c = graphs.autoencoder_forward(w, crop)
    assert np.abs(want_c[:, m:-m, m:-m] - inner_full).max() < 1e-2


def test_roi_sweep_1024_matches_oracle(cic, precision):
    """BASELINE configs[2]: 1024x1024 image = 16 tiles of 256x256 through the adaptive model at three target rates,
    plus the hq_ratio sweep over np.linspace(0.1, 2.0, 10) (GAN_test.py:543)."""
    models, ws = _adaptive(cic, (256, 256, 3), 512)
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(3, 1024, 1024, seed=62))
    mask = cic.synth.synth_masks(3, 1024, 1024, seed=62)
    bpp = np.array([[0.1], [1.0], [2.0]], np.float32)
    out, want, ex = _check_adaptive(cic, precision, models, ws, img, mask, bpp, 256)
    levels = cic.synth.rate_control_bpps().astype(np.float32)
    sweep = cic.ops.hq_ratio_sweep(mask, levels).cpu().numpy()
    assert np.all(np.diff(sweep, axis=1) > 0)
    for i, b in enumerate((0.1, 1.0, 2.0)):                                    # the sweep and the full model agree
        k = int(np.argmin(np.abs(levels - b)))
        if abs(levels[k] - b) < 1e-6:
            assert abs(sweep[i, k] - out["hq_ratio_sum"][i] / (1024 * 1024)) < 1e-6
    acc = cic.gan.bpp_accounting(float(out["hq_ratio_sum"][1] / (1024 * 1024)))
    assert 0.25 <= acc["actual_bpp"] <= 0.5
