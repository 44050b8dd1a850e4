"""GPU parity of the model graphs (autoencoder, encoder, generator, adaptive codec) against the CPU
oracle, through the drop-in Python interface -> C ABI.

North-star criteria: quantised symbols bit-exact except where the oracle's fp32 pre-round value lies
within 1e-3 of a rounding boundary (mismatch count reported); reconstruction within 1e-2 max-abs in
[0,1] pixel space; PSNR within 0.05 dB.  The fp32 arithmetic mode is held to much tighter bounds.
"""
import os

import numpy as np
import pytest
import torch

from oracle import graphs, metrics

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

PRECISIONS = ["fp32", "tc"]
# max-abs bound in [0,1] pixel space / latent bound, per arithmetic mode
RECON_TOL = {"fp32": 2e-5, "tc": 1e-2}
LATENT_TOL = {"fp32": 2e-4, "tc": 2e-3}


@pytest.fixture(params=PRECISIONS)
def precision(request, cic):
    if request.param == "tc" and os.environ.get("CIC_SKIP_TC") == "1":
        pytest.skip("CIC_SKIP_TC=1")
    old = cic.get_precision()
    cic.set_precision(request.param)
    yield request.param
    cic.set_precision(old)


def symbol_parity(got_sym, want_sym, want_pre):
    """Returns (#mismatches, #mismatches outside the 1e-3 boundary band)."""
    bad = np.asarray(got_sym).astype(np.int64) != np.asarray(want_sym).astype(np.int64)
    frac = np.abs(want_pre - np.floor(want_pre))
    near = np.abs(frac - 0.5) < 1e-3
    return int(bad.sum()), int((bad & ~near).sum())


@pytest.mark.parametrize("B,H,W", [(2, 32, 48), (1, 256, 256), (3, 64, 64), (1, 120, 68)])
def test_autoencoder_matches_oracle(cic, precision, B, H, W):
    import train_autoencoder as tr
    model = tr.build_autoencoder((H, W, 3))
    w = cic.weights.synthetic_autoencoder(seed=42)
    model.set_weights_dict(w)
    x = cic.synth.to_unit_range(cic.synth.synth_images_u8(B, H, W, seed=43))
    y = model.predict(x)
    want = graphs.autoencoder_forward(w, x)
    assert y.shape == want.shape and y.dtype == np.float32
    err = np.abs(y - want).max()
    assert err < RECON_TOL[precision], f"max-abs {err}"
    # batch-1 loop of test_autoencoder.py:83-85 gives the same pixels as one batch
    # (fp32: identical.  tc: the tile geometry of the tensor-core layers follows the batch size, so fp32 partial sums are grouped
    # differently and a bf16 activation can round the other way - isolated 3x3 patches of pixels move by a few 1e-4.)
    y0 = model.predict(np.expand_dims(x[0], axis=0))[0]
    np.testing.assert_allclose(y0, y[0], atol=1e-6 if precision == "fp32" else 2e-3)
    assert np.mean(np.abs(y0 - y[0]) > 1e-6) < 0.15
    # uint8 "quantiser": truncation; 1-LSB flips only where y*255 sits next to an integer
    r = cic.autoencoder.evaluate_batch(model, x)
    y8 = r["compressed_u8"].cpu().numpy()
    want8 = graphs.autoencoder_output_u8(want)
    diff = np.abs(y8.astype(int) - want8.astype(int))
    assert diff.max() <= 1
    x8 = (x * 255).astype(np.uint8)
    for i in range(B):
        assert abs(r["psnr"][i] - metrics.ae_calculate_psnr(x8[i], y8[i])) < 1e-9     # metrics on OUR pixels: exact
        assert abs(r["ssim"][i] - metrics.ae_calculate_ssim(x8[i], y8[i])) < 1e-9
        assert abs(r["psnr"][i] - metrics.ae_calculate_psnr(x8[i], want8[i])) < 0.05    # vs the oracle's pixels
        assert abs(r["mse"][i] - metrics.ae_calculate_mse(x8[i], y8[i])) < 1e-9


def test_autoencoder_golden_fixture(cic, precision):
    g = np.load(os.path.join(HERE, "golden", "oracle_small.npz"))
    import train_autoencoder as tr
    model = tr.build_autoencoder((32, 48, 3))
    model.set_weights_dict(cic.weights.synthetic_autoencoder(seed=42))
    y = model.predict(cic.synth.to_unit_range(cic.synth.synth_images_u8(2, 32, 48, seed=43)))
    assert np.abs(y - g["ae_y"]).max() < RECON_TOL[precision]


def test_autoencoder_rejects_bad_shapes(cic):
    import train_autoencoder as tr
    model = tr.build_autoencoder((32, 32, 3))
    with pytest.raises(ValueError):
        model.predict(np.zeros((1, 30, 32, 3), np.float32))
    with pytest.raises(ValueError):
        model.predict(np.zeros((1, 32, 32, 4), np.float32))
    assert model.predict(np.zeros((0, 32, 32, 3), np.float32)).shape == (0, 32, 32, 3)


def _adaptive(cic, img_shape, base, seed=42):
    import GAN_functions as gf
    models = gf.build_adaptive_compression_model(img_shape, base, target_bpp=True)
    ws = cic.weights.synthetic_adaptive(img_shape, base, seed=seed)
    models["adaptive_model"].set_weights_dict(ws)
    return models, ws


def _check_adaptive(cic, precision, models, ws, img, mask, bpp, tile):
    am = models["adaptive_model"]
    out = am.forward_device([cic.runtime.to_device_f32(img), cic.runtime.to_device_f32(mask),
                             cic.runtime.to_device_f32(bpp)], extras=True)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    # oracle per tile
    n, H, W, _ = img.shape
    ty, tx = H // tile, W // tile
    tiles_i = img.reshape(n, ty, tile, tx, tile, 3).transpose(0, 1, 3, 2, 4, 5).reshape(-1, tile, tile, 3)
    tiles_m = mask.reshape(n, ty, tile, tx, tile, 1).transpose(0, 1, 3, 2, 4, 5).reshape(-1, tile, tile, 1)
    bpp_t = np.repeat(bpp.reshape(-1), ty * tx).reshape(-1, 1)
    want, ex = graphs.adaptive_forward(ws, tiles_i, tiles_m, bpp_t, return_extras=True)
    untile = lambda t, c: t.reshape(n, ty, tx, tile, tile, c).transpose(0, 1, 3, 2, 4, 5).reshape(n, H, W, c)
    # latents and symbols
    assert np.abs(out["hq_latent"] - ex["hq_latent"]).max() < LATENT_TOL[precision]
    assert np.abs(out["lq_latent"] - ex["lq_latent"]).max() < LATENT_TOL[precision]
    report = {}
    for br in ("hq", "lq"):
        nbad, nout = symbol_parity(out[f"{br}_symbols"], ex[f"{br}_sym"], ex[f"{br}_pre"])
        report[br] = (nbad, nout, out[f"{br}_symbols"].size)
        assert nout == 0, f"{br}: {nout} symbol mismatches outside the 1e-3 boundary band (of {nbad} total)"
        np.testing.assert_allclose(out[f"{br}_scale"], ex[f"{br}_scale"].ravel(), rtol=2e-5)
    print(f"[{precision}] symbol mismatches (total, outside band, n): {report}")
    # dequantised latents agree wherever the symbols agree
    ok = out["hq_symbols"] == ex["hq_sym"].astype(np.int32)
    np.testing.assert_allclose(out["hq_latent_q"][ok], want[1][ok], rtol=3e-5, atol=1e-6)
    # reconstructions in [0,1] pixel space: un-blended generators and the blended output.
    # Tiles whose symbols differ inside the band decode differently by construction; compare those
    # only on the blend identity below.
    same_tiles = np.array([np.array_equal(out["hq_symbols"][t], ex["hq_sym"][t].astype(np.int32)) and
                           np.array_equal(out["lq_symbols"][t], ex["lq_sym"][t].astype(np.int32)) for t in range(n * ty * tx)])
    hq_t = out["hq_out"].reshape(n, ty, tile, tx, tile, 3).transpose(0, 1, 3, 2, 4, 5).reshape(-1, tile, tile, 3)
    bl_t = out["blended"].reshape(n, ty, tile, tx, tile, 3).transpose(0, 1, 3, 2, 4, 5).reshape(-1, tile, tile, 3)
    assert same_tiles.any()
    assert np.abs(hq_t[same_tiles] - ex["hq_out"][same_tiles]).max() / 2 < RECON_TOL[precision]
    assert np.abs(bl_t[same_tiles] - want[0][same_tiles]).max() / 2 < RECON_TOL[precision]
    np.testing.assert_allclose(out["dt"], untile(want[4], 1), atol=3e-6)
    np.testing.assert_allclose(out["rd_params"], want[3], atol=2e-5)
    # blend identity on our own tensors
    np.testing.assert_allclose(out["blended"], out["hq_out"] * out["dt"] + out["lq_out"] * (1 - out["dt"]), atol=2e-6)
    # hq_ratio / bpp accounting (GAN_test.py:310-325)
    hq_ratio = out["hq_ratio_sum"] / (H * W)
    np.testing.assert_allclose(hq_ratio, untile(want[4], 1).reshape(n, -1).mean(1, dtype=np.float64), atol=1e-5)
    # PSNR within 0.05 dB on tiles that decoded the same symbols
    for t in np.flatnonzero(same_tiles)[:4]:
        a = metrics.compute_metrics(tiles_i[t], want[0][t])
        b = metrics.compute_metrics(tiles_i[t], bl_t[t])
        assert abs(a["psnr"] - b["psnr"]) < 0.05 and abs(a["ssim"] - b["ssim"]) < 1e-3
    return out, want, ex


def test_adaptive_small_matches_oracle(cic, precision, small_cfg):
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(3, 64, 64, seed=44))
    mask = cic.synth.synth_masks(3, 64, 64, seed=44)
    bpp = np.array([[0.1], [1.0], [2.0]], np.float32)
    out, want, ex = _check_adaptive(cic, precision, models, ws, img, mask, bpp, 64)
    g = np.load(os.path.join(HERE, "golden", "oracle_small.npz"))               # committed fixture
    assert np.abs(out["hq_latent"] - g["ad_hq_latent"]).max() < LATENT_TOL[precision]
    np.testing.assert_allclose(out["dt"], g["ad_dt"], atol=3e-6)


def test_adaptive_reference_size_matches_oracle(cic, precision):
    """The reference configuration: 256x256 tiles, base latent 512 (HQ 1024 + attention over 1024 tokens)."""
    models, ws = _adaptive(cic, (256, 256, 3), 512)
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(2, 256, 256, seed=45))
    mask = cic.synth.synth_masks(2, 256, 256, seed=45)
    bpp = np.array([[0.1], [1.0]], np.float32)
    _check_adaptive(cic, precision, models, ws, img, mask, bpp, 256)


def test_adaptive_tiled_image_equals_tiles(cic, precision, small_cfg):
    """Images larger than the model tile are coded as independent tiles (BASELINE configs 2-5)."""
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(2, 128, 192, seed=46))
    mask = cic.synth.synth_masks(2, 128, 192, seed=46)
    bpp = np.array([[0.5], [1.5]], np.float32)
    _check_adaptive(cic, precision, models, ws, img, mask, bpp, 64)


def test_submodels_match_oracle(cic, precision, small_cfg):
    import GAN_functions as gf
    shape, base = small_cfg["img_shape"], small_cfg["base"]
    ws = cic.weights.synthetic_adaptive(shape, base, seed=5)
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(2, 64, 64, seed=47))
    enc = gf.build_encoder(shape, 2 * base, name="hq_encoder", add_attention=True)
    enc.set_weights_dict(ws["hq_encoder"])
    got = enc.predict(img)
    want = graphs.encoder_forward(ws["hq_encoder"], img, True)
    assert [g.shape for g in got] == [w.shape for w in want]
    assert np.abs(got[0] - want[0]).max() < LATENT_TOL[precision]
    for g, w in zip(got[1:], want[1:]):
        assert np.abs(g - w).max() < (1e-4 if precision == "fp32" else 2e-2)
    gen = gf.build_generator(2 * base, shape, name="hq_generator")
    gen.set_weights_dict(ws["hq_generator"])
    y = gen.predict(want)                                   # oracle latents + skips in, like the reference's list input
    wy = graphs.generator_forward(ws["hq_generator"], *want)
    assert np.abs(y - wy).max() / 2 < RECON_TOL[precision]
    # __call__(list, training=False)[i].numpy() protocol of GAN_functions.py:867-871
    outs = enc([img], training=False)
    assert np.array_equal(outs[0].numpy(), got[0]) and outs[1][0].numpy().shape == (32, 32, 64)
    sal = gf.build_latent_saliency_model(2 * base)
    sal.set_weights_dict(ws["latent_saliency_hq"])
    np.testing.assert_allclose(sal.predict(want[0]), graphs.latent_saliency_forward(ws["latent_saliency_hq"], want[0]), atol=1e-5)
    rd = gf.build_rate_distortion_optimizer(shape, None)
    rd.set_weights_dict(ws["rd_optimizer"])
    mask = cic.synth.synth_masks(2, 64, 64, seed=47)
    bpp = np.array([[0.3], [1.7]], np.float32)
    np.testing.assert_allclose(rd.predict([img, mask, bpp]), graphs.rd_optimizer_forward(ws["rd_optimizer"], mask, bpp), atol=2e-5)
    q = gf.AdaptiveQuantizationLayer()([want[0], np.array([[0.4], [0.6]], np.float32), np.array([[0.7], [0.2]], np.float32)])
    wq = graphs.adaptive_quantize(want[0], np.array([[0.4], [0.6]], np.float32), np.array([[0.7], [0.2]], np.float32))[0]
    assert np.mean(np.abs(q.numpy() - wq) > 1e-5) < 0.01


def test_compress_and_reconstruct_and_rate_control(cic, precision, small_cfg):
    import GAN_test as gt
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(4, 64, 64, seed=48))
    mask = cic.synth.synth_masks(4, 64, 64, seed=48)[..., 0]
    r = gt.compress_and_reconstruct(img[0], models, target_bpp=1.0, mask=mask[0])
    assert set(r) == {"saliency_map", "compressed_img", "hq_latent", "lq_latent", "rd_params", "bit_allocation", "metrics",
                      "compression_ratio", "actual_bpp", "target_bpp", "hq_ratio", "lq_ratio"}
    assert r["compressed_img"].shape == (64, 64, 3) and r["hq_latent"].shape == (64,)
    assert r["actual_bpp"] == pytest.approx((r["hq_ratio"] * 64 + r["lq_ratio"] * 32) * 32 / 4096)
    want = graphs.adaptive_forward(ws, img[:1], mask[:1, :, :, None], np.array([[1.0]], np.float32))
    wm = metrics.compute_metrics(img[0], want[0][0])
    assert abs(r["hq_ratio"] - want[4].mean()) < 1e-5
    if np.array_equal(np.rint(r["hq_latent"] * 1e3), np.rint(want[1][0] * 1e3)):
        assert abs(r["metrics"]["psnr"] - wm["psnr"]) < 0.05
    names = [f"im{i}" for i in range(4)]
    fast = gt.test_rate_control(models, list(img), names, masks=list(mask))
    full = gt.test_rate_control(models, list(img), names, masks=list(mask), full_model=True)
    assert len(fast["hq_ratio"]) == 40 and fast["target_bpp"][:2] == pytest.approx([0.1, 0.1 + 1.9 / 9])
    np.testing.assert_allclose(fast["hq_ratio"], full["hq_ratio"], atol=1e-6)    # one sweep pass == 10 full predictions
    np.testing.assert_allclose(fast["actual_bpp"], [(h * 64 + (1 - h) * 32) * 32 / 4096 for h in fast["hq_ratio"]], rtol=1e-12)
    per_img = np.array(fast["hq_ratio"]).reshape(4, 10)
    assert np.all(np.diff(per_img, axis=1) > 0)
    # no mask: the reference's whole front end (compute_saliency_map 'combined' -> create_saliency_mask, GAN_test.py:279-280) on the GPU
    from oracle import saliency as osal
    r2 = gt.compress_and_reconstruct(img[0], models, target_bpp=1.0)
    want_mask = osal.create_saliency_mask(osal.compute_saliency_map(img[0], "combined"), smooth=True)   # oracle map, cv2 mask
    np.testing.assert_allclose(r2["saliency_map"], want_mask, atol=1e-4)
    want2 = graphs.adaptive_forward(ws, img[:1], want_mask[None, :, :, None], np.array([[1.0]], np.float32))
    assert abs(r2["hq_ratio"] - want2[4].mean()) < 1e-4


def test_pipelined_predict_equals_predict(cic, precision, small_cfg):
    """predict_pipelined (chunks over copy-in / compute / copy-out streams) returns what predict returns."""
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    am = models["adaptive_model"]
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(5, 128, 64, seed=49))
    mask = cic.synth.synth_masks(5, 128, 64, seed=49)
    bpp = np.linspace(0.2, 1.8, 5, dtype=np.float32).reshape(5, 1)
    want = am.predict([img, mask, bpp])
    seen = []
    got, extra = am.predict_pipelined([img, mask, bpp], n_chunks=3,
                                      on_chunk=lambda d_in, outs: seen.append((d_in[0].shape[0], outs["hq_ratio_sum"].clone())))
    assert [n for n, _ in seen] == [1, 2, 2] and len(extra) == 3
    for g, w in zip(got, want):
        assert g.shape == w.shape
        np.testing.assert_allclose(g, w, atol=0 if precision == "fp32" else 2e-2, rtol=0)
    ratios = torch.cat([r for _, r in seen]).cpu().numpy() / (128 * 64)
    np.testing.assert_allclose(ratios, want[4].reshape(5, -1).mean(1, dtype=np.float64), atol=1e-6)
    import train_autoencoder as tr
    ae = tr.build_autoencoder((32, 32, 3))
    ae.set_weights_dict(cic.weights.synthetic_autoencoder(seed=42))
    x = cic.synth.to_unit_range(cic.synth.synth_images_u8(7, 32, 32, seed=50))
    y, _ = ae.predict_pipelined(x, n_chunks=4)
    # fp32: bit-identical.  tc: the tile shape (hence the tap accumulation order) may change with the chunk's batch size,
    # so an intermediate bf16 rounding can flip by one ulp
    np.testing.assert_allclose(y, ae.predict(x), atol=0 if precision == "fp32" else 1e-3, rtol=0)


def test_phased_predict_equals_predict(cic, precision, small_cfg):
    """predict_phased (encoder convs per upload chunk, Dense / quantiser once per batch, decoders per download chunk; eager on
    the first call, CUDA graphs from the second) returns what predict returns."""
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    am = models["adaptive_model"]
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(7, 128, 64, seed=51))
    mask = cic.synth.synth_masks(7, 128, 64, seed=51)
    bpp = np.linspace(0.2, 1.8, 7, dtype=np.float32).reshape(7, 1)
    if precision == "fp32":
        with pytest.raises(cic._lib.CicError, match="tensor-core plans only"):
            am.predict_phased([img, mask, bpp])
        return
    want = am.predict([img, mask, bpp])
    on_chunk = lambda d_in, outs: (d_in[0].shape[0], outs["hq_ratio_sum"].clone())  # noqa: E731
    for call in range(3):                                                       # eager, graph capture, graph replay
        got, extra = am.predict_phased([img, mask, bpp], enc_chunks=[1, 2, 4], dec_chunks=[4, 2, 1], on_chunk=on_chunk)
        assert [n for n, _ in extra] == [4, 2, 1]
        for g, w in zip(got, want):
            assert g.shape == w.shape
            np.testing.assert_allclose(g, w, atol=2e-2, rtol=0)
        # quantised latents: kernel selection can change with a chunk's tile count (a bf16 ulp upstream), which may move a
        # pre-round value across a rounding boundary
        assert np.mean(got[1] != want[1]) < 0.01
        ratios = torch.cat([r for _, r in extra]).cpu().numpy() / (128 * 64)
        np.testing.assert_allclose(ratios, want[4].reshape(7, -1).mean(1, dtype=np.float64), atol=1e-6)
    got, _ = am.predict_phased([img, mask, bpp])                                 # default chunking (n < 8: one chunk)
    np.testing.assert_allclose(got[0], want[0], atol=2e-2, rtol=0)
    with pytest.raises(ValueError, match="add up"):
        am.predict_phased([img, mask, bpp], enc_chunks=[3, 3])
    # uint8 image in, uint8 blended out: the reference's load / save conventions applied on the device
    img_u8 = cic.synth.synth_images_u8(7, 128, 64, seed=51)
    for call in range(3):
        got8, _ = am.predict_phased([img_u8, mask, bpp], enc_chunks=[1, 2, 4], dec_chunks=[4, 2, 1], u8_io=True)
        assert got8[0].dtype == np.uint8
        want8 = ((want[0] + 1) * np.float32(127.5)).astype(np.uint8)
        assert np.abs(got8[0].astype(int) - want8.astype(int)).max() <= 3          # 2e-2 in [-1,1] = 2.55 uint8 steps
        np.testing.assert_allclose(got8[4], want[4], atol=1e-6)
    with pytest.raises(ValueError, match="uint8"):
        am.predict_phased([img, mask, bpp], u8_io=True)


def test_linearity_of_blend_at_full_size(cic):
    """Size-independent property at a BASELINE-scale shape (1024x1024): blend(hq, hq) == hq and
    blend is affine in (hq, lq)."""
    rng = np.random.default_rng(0)
    dev = "cuda"
    hq = torch.rand((2, 1024, 1024, 3), device=dev) * 2 - 1
    lq = torch.rand((2, 1024, 1024, 3), device=dev) * 2 - 1
    mask = torch.from_numpy(cic.synth.synth_masks(2, 1024, 1024)).to(dev)
    bpp = torch.tensor([0.4, 1.6], device=dev)
    same, dt, s = cic.ops.roi_mask_blend(hq, hq, mask, bpp)
    assert (same - hq).abs().max().item() < 2e-7
    out, _, _ = cic.ops.roi_mask_blend(hq, lq, mask, bpp)
    assert (out - (hq * dt + lq * (1 - dt))).abs().max().item() < 1e-6
    assert torch.allclose(s, dt.double().sum(dim=(1, 2, 3)), rtol=1e-9)
    sweep = cic.ops.hq_ratio_sweep(mask, bpp)
    assert abs(sweep[0, 0].item() - s[0].item() / (1024 * 1024)) < 1e-7


def test_stream_predict_equals_predict(cic, small_cfg):
    """predict_stream (batches overlapping each other on three streams, one CUDA graph per slot from its second use) yields, in
    order, what predict returns for every batch - float32 and uint8 wire formats, with and without the dt map."""
    cic.set_precision("tc")
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    am = models["adaptive_model"]
    batches, wants, batches8 = [], [], []
    for k in range(6):
        img_u8 = cic.synth.synth_images_u8(3, 128, 64, seed=200 + k)
        img = cic.synth.to_signed_range(img_u8)
        mask = cic.synth.synth_masks(3, 128, 64, seed=200 + k)
        bpp = np.linspace(0.2 + 0.1 * k, 1.8, 3, dtype=np.float32).reshape(3, 1)
        batches.append([img, mask, bpp])
        batches8.append([img_u8, mask, bpp])
        wants.append(am.predict([img, mask, bpp]))
    on_batch = lambda d_in, outs: outs["hq_ratio_sum"].clone()  # noqa: E731
    got = [([np.array(o) for o in outs], ex.cpu().numpy().copy()) for outs, ex in am.predict_stream(iter(batches), on_batch=on_batch)]
    assert len(got) == 6
    for (outs, ratio), want in zip(got, wants):
        for g, w in zip(outs, want):
            assert g.shape == w.shape
            np.testing.assert_allclose(g, w, atol=2e-2, rtol=0)
        assert np.mean(outs[1] != want[1]) < 0.01
        np.testing.assert_allclose(ratio / (128 * 64), want[4].reshape(3, -1).mean(1, dtype=np.float64), atol=1e-6)
    got8 = [[np.array(o) for o in outs] for outs, _ in am.predict_stream(iter(batches8), u8_io=True, want_dt=False, depth=3)]
    for outs, want in zip(got8, wants):
        assert outs[0].dtype == np.uint8
        want8 = ((want[0] + 1) * np.float32(127.5)).astype(np.uint8)
        assert np.abs(outs[0].astype(int) - want8.astype(int)).max() <= 3
        np.testing.assert_allclose(outs[4], want[4].reshape(3, -1).mean(1, dtype=np.float64), atol=1e-6)
    assert list(am.predict_stream(iter([]))) == []
    # jpeg_out: the reconstruction leaves the device as the file save_image would write (GAN_functions.py:41-50) - equal to OpenCV's
    # encoding of the uint8 reconstruction the u8_io stream returns for the same batch (same kernels, same graph inputs)
    import cv2
    gotj = [[np.array(o) for o in outs] for outs, _ in am.predict_stream(iter(batches8), u8_io=True, want_dt=False, depth=3, jpeg_out=2.0)]
    assert len(gotj) == 6
    for outs, ref8 in zip(gotj, got8):
        assert len(outs) == 6 and outs[0].dtype == np.uint8 and outs[5].shape == (3,)
        for i in range(3):
            size = int(outs[5][i])
            assert 623 < size <= outs[0].shape[1]
            want = bytes(cv2.imencode(".jpg", cv2.cvtColor(ref8[0][i], cv2.COLOR_RGB2BGR))[1])
            assert outs[0][i, :size].tobytes() == want
        np.testing.assert_array_equal(outs[1], ref8[1])
    with pytest.raises(ValueError, match="u8_io"):
        list(am.predict_stream(iter(batches), jpeg_out=1.0))


def test_rate_sweep_equals_one_predict_per_level(cic, small_cfg):
    """rate_sweep_device (BASELINE configs[2]: encoders once, quantiser + generators + blend per target bpp) gives, per level, what
    the reference's loop of full predictions gives (GAN_test.py:565-573): hq_ratio, blended image, quantised latents."""
    cic.set_precision("tc")
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    am = models["adaptive_model"]
    img = cic.synth.to_signed_range(cic.synth.synth_images_u8(2, 128, 192, seed=52))
    mask = cic.synth.synth_masks(2, 128, 192, seed=52)
    levels = [0.1, 0.944444, 2.0]
    seen = {}

    def on_level(k, ins, outs):
        seen[k] = (outs["blended"].cpu().numpy().copy(), outs["hq_latent_q"].cpu().numpy().copy(), outs["dt"].cpu().numpy().copy())
    ratios = am.rate_sweep_device(cic.runtime.to_device_f32(img), cic.runtime.to_device_f32(mask), levels, on_level).cpu().numpy()
    assert ratios.shape == (3, 2) and np.all(np.diff(ratios, axis=0) > 0)
    for k, lv in enumerate(levels):
        want = am.predict([img, mask, np.full((2, 1), lv, np.float32)])
        np.testing.assert_allclose(ratios[k], want[4].reshape(2, -1).mean(1, dtype=np.float64), atol=1e-6)
        np.testing.assert_allclose(seen[k][2], want[4], atol=1e-6)
        np.testing.assert_allclose(seen[k][0], want[0], atol=2e-2, rtol=0)
        assert np.mean(seen[k][1] != want[1]) < 0.01
    sweep = cic.ops.hq_ratio_sweep(mask, np.array(levels, np.float32)).cpu().numpy()          # the mask-only fast path agrees
    np.testing.assert_allclose(sweep.T, ratios, atol=1e-6)


def test_host_paths_on_ragged_sizes(cic, small_cfg):
    """predict_phased / predict_stream / rate_sweep_device on an image size that is not a multiple of the tile: the same cropped
    outputs as predict (whose ragged-size parity against the oracle is test_adaptive_ragged_small)."""
    cic.set_precision("tc")
    models, ws = _adaptive(cic, small_cfg["img_shape"], small_cfg["base"])
    am = models["adaptive_model"]
    n, h, w = 5, 100, 148
    img_u8 = cic.synth.synth_images_u8(n, h, w, seed=53)
    img = cic.synth.to_signed_range(img_u8)
    mask = cic.synth.synth_masks(n, h, w, seed=53)
    bpp = np.linspace(0.3, 1.7, n, dtype=np.float32).reshape(n, 1)
    want = am.predict([img, mask, bpp])
    assert want[0].shape == (n, h, w, 3) and want[1].shape == (n * 2 * 3, 2 * small_cfg["base"])
    for call in range(3):
        got, _ = am.predict_phased([img, mask, bpp], enc_chunks=[2, 3], dec_chunks=[3, 2])
        for g, wv in zip(got, want):
            assert g.shape == wv.shape
            np.testing.assert_allclose(g, wv, atol=2e-2, rtol=0)
    outs = [([np.array(o) for o in o5]) for o5, _ in am.predict_stream(iter([[img_u8, mask, bpp]] * 3), u8_io=True)]
    want8 = ((want[0] + 1) * np.float32(127.5)).astype(np.uint8)
    for o5 in outs:
        assert o5[0].shape == (n, h, w, 3) and np.abs(o5[0].astype(int) - want8.astype(int)).max() <= 3
        np.testing.assert_allclose(o5[4], want[4], atol=1e-6)
    ratios = am.rate_sweep_device(cic.runtime.to_device_f32(img), cic.runtime.to_device_f32(mask), [1.0]).cpu().numpy()
    full = am.predict([img, mask, np.ones((n, 1), np.float32)])
    np.testing.assert_allclose(ratios[0], full[4].reshape(n, -1).mean(1, dtype=np.float64), atol=1e-6)
