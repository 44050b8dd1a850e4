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


# ---- the tiled codec at sizes that are not multiples of the 256 tile (BASELINE configs[3], [4]) ---------------------------------
def _tiled_parity(cic, am, ws, img, mask, bpp, tile, sample, precision, seed=0):
    """forward_device on the whole batch, the oracle on `sample` of its tiles (edge-padded like the product pads), the three
    north-star criteria + dt + hq_ratio."""
    from oracle import parity, tiling
    n, h, w, _ = img.shape
    out = am.forward_device([cic.runtime.to_device_f32(img), cic.runtime.to_device_f32(mask), cic.runtime.to_device_f32(bpp)],
                            extras=True)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    ty, tx = tiling.grid(h, w, tile)
    assert got["hq_symbols"].shape[0] == n * ty * tx and got["blended"].shape == (n, h, w, 3)
    ref = tiling.adaptive_forward_tiled(ws, img, mask, bpp, tile, tiles=tiling.sample_tiles(n * ty * tx, sample, seed))
    rep = parity.adaptive_parity(got, ref, n, h, w, tile)
    print(f"[{precision}] {n} x {h}x{w}: {rep}")
    parity.assert_north_star(rep, recon_tol=2e-5 if precision == "fp32" else 1e-2)
    np.testing.assert_allclose(got["hq_ratio_sum"] / (h * w), tiling.hq_ratio(mask, bpp), atol=1e-5)     # unpadded pixels only
    np.testing.assert_allclose(got["rd_params"][ref["tiles"]], ref["rd_params"], atol=2e-5)
    return got, ref, rep


def test_adaptive_ragged_small(cic, precision, small_cfg):
    """Sizes that are not multiples of the model tile, including an image smaller than one tile and odd sizes."""
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    am = models["adaptive_model"]
    for (n, h, w) in ((2, 100, 150), (1, 40, 64), (3, 65, 63)):
        img = cic.synth.to_signed_range(cic.synth.synth_images_u8(n, h, w, seed=70 + h))
        mask = cic.synth.synth_masks(n, h, w, seed=70 + h)
        bpp = np.linspace(0.2, 1.9, n, dtype=np.float32).reshape(n, 1)
        got, ref, rep = _tiled_parity(cic, am, ws, img, mask, bpp, 64, 10**6, precision)
        # the host API returns the same thing, cropped arrays included
        outs = am.predict([img, mask, bpp])
        assert outs[0].shape == (n, h, w, 3) and outs[4].shape == (n, h, w, 1)
        np.testing.assert_array_equal(outs[0], got["blended"])


def test_adaptive_1080p_frame(cic, precision):
    """BASELINE configs[3]: one 1920x1080 frame = 5 x 8 = 40 tiles of 256x256 (the last tile row covers 56 image rows)."""
    models, ws = _adaptive(cic, (256, 256, 3), 512)
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(1, 1080, 1920, seed=63))
    mask = cic.synth.synth_masks(1, 1080, 1920, seed=63)
    bpp = np.array([[1.0]], np.float32)
    _tiled_parity(cic, models["adaptive_model"], ws, img, mask, bpp, 256, 40 if precision == "tc" else 12, precision)


def test_adaptive_4k_frame(cic):
    """BASELINE configs[4]: one 3840x2160 image = 9 x 15 = 135 tiles; the oracle runs on 12 sampled tiles (first, last = the
    ragged corner, 10 in between)."""
    cic.set_precision("tc")
    models, ws = _adaptive(cic, (256, 256, 3), 512)
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(1, 2160, 3840, seed=64))
    mask = cic.synth.synth_masks(1, 2160, 3840, seed=64)
    bpp = np.array([[2.0]], np.float32)
    _tiled_parity(cic, models["adaptive_model"], ws, img, mask, bpp, 256, 12, "tc", seed=3)


def test_predict_phased_bench_shape_matches_oracle(cic):
    """The API bench.py's `e2e` times, at the bench shape (64 x 512x512, default chunks, CUDA-graph replay on the third call),
    against the ORACLE on 12 sampled tiles - float32 and uint8 image I/O."""
    from oracle import parity, tiling
    cic.set_precision("tc")
    models, ws = _adaptive(cic, (256, 256, 3), 512)
    am = models["adaptive_model"]
    n, hw = 64, 512
    img_u8 = cic.synth.synth_images_u8(n, hw, hw, seed=cic.synth.SEED_BASE + 1)
    img = cic.synth.to_signed_range(img_u8)
    mask = cic.synth.synth_masks(n, hw, hw, seed=cic.synth.SEED_BASE + 1)
    bpp = np.full((n, 1), 1.0, np.float32)
    sel = tiling.sample_tiles(n * 4, 12, seed=1)
    ref = tiling.adaptive_forward_tiled(ws, img, mask, bpp, 256, tiles=sel)
    pins = [torch.from_numpy(a).pin_memory() for a in (img, mask, bpp)]
    for call in range(3):                                              # eager, capture, replay
        outs, _ = am.predict_phased(pins)
    blended, hq_q, lq_q, rd, dt = [np.array(o) for o in outs]
    # symbols are not an output of the reference API: recover them from the dequantised latents with the oracle's scales
    got = {"blended": blended, "dt": dt}
    full_hq = np.zeros((n * 4, hq_q.shape[1]), np.int64)
    full_lq = np.zeros((n * 4, lq_q.shape[1]), np.int64)
    full_hq[sel] = np.rint(hq_q[sel].astype(np.float64) * ref["hq_scale"].reshape(-1, 1))
    full_lq[sel] = np.rint(lq_q[sel].astype(np.float64) * ref["lq_scale"].reshape(-1, 1))
    got["hq_symbols"], got["lq_symbols"] = full_hq, full_lq
    rep = parity.adaptive_parity(got, ref, n, hw, hw, 256)
    print(f"predict_phased at the bench shape vs oracle: {rep}")
    parity.assert_north_star(rep)
    np.testing.assert_allclose(rd[sel], ref["rd_params"], atol=2e-5)
    # uint8 wire format: same latents, blended within one uint8 step of the oracle's save_image() bytes on identical-symbol tiles
    pins8 = [torch.from_numpy(img_u8).pin_memory(), pins[1], pins[2]]
    for call in range(3):
        outs8, _ = am.predict_phased(pins8, u8_io=True, want_dt=False)
    assert outs8[0].dtype == np.uint8
    np.testing.assert_array_equal(outs8[1], hq_q)
    np.testing.assert_allclose(outs8[4], tiling.hq_ratio(mask, bpp), atol=1e-5)
    worst = 0
    for k, t in enumerate(sel):
        i, y0, x0, vh, vw = tiling.tile_window(int(t), hw, hw, 256)
        if not (np.array_equal(full_hq[t], ref["hq_sym"][k].astype(np.int64)) and np.array_equal(full_lq[t], ref["lq_sym"][k].astype(np.int64))):
            continue
        want8 = ((ref["blended"][k] + 1) * np.float32(127.5)).astype(np.uint8)
        worst = max(worst, int(np.abs(outs8[0][i, y0:y0 + vh, x0:x0 + vw].astype(int) - want8.astype(int)).max()))
    assert worst <= 3, worst                                            # 1e-2 in [0,1] = 2.55 uint8 steps
