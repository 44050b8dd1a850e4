"""GPU parity of the stand-alone operators against the CPU oracle, through the C ABI (ctypes)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import graphs, metrics

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))


def test_library_sees_b200(cic):
    import ctypes
    sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    cic._lib.check(cic._lib.lib.cic_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr)))
    assert maj.value == 10 and sm.value >= 100


def test_rate_scalars_bit_exact(cic):
    bpps = np.concatenate([np.linspace(0.1, 2.0, 10), [0.0, -1.0, 5.0, 7.5, 0.3111111]]).astype(np.float32)
    t, thr, qs = (v.cpu().numpy() for v in cic.ops.rate_scalars(bpps))
    ot, othr, oqs = (v.numpy().ravel() for v in graphs.rate_scalars(bpps))
    np.testing.assert_array_equal(t, ot)
    np.testing.assert_array_equal(thr, othr)
    np.testing.assert_array_equal(qs, oqs)
    for row in KA["rate_scalars"]:
        _, thr1, qs1 = cic.ops.rate_scalars([row["bpp"]])
        assert abs(thr1.item() - row["thr"]) < 2e-6 and abs(qs1.item() - row["qs"]) < 2e-6


def test_quantizer_known_answers(cic):
    for row in KA["quantizer"]:
        _, _, qs = cic.ops.rate_scalars([row["bpp"]])
        r = cic.ops.quantize_latent(np.array([[row["latent"]]], np.float32), [row["sal"]], qs, want=("deq", "symbols", "pre", "scale"))
        assert int(r["symbols"].item()) == row["symbol"]                        # includes the half-to-even cases
        assert abs(r["scale"].item() - row["scale"]) < 1e-5 * max(1, row["scale"])
        assert abs(r["deq"].item() - row["deq"]) < 2e-6 * max(1, abs(row["deq"]))


@pytest.mark.parametrize("batch,L", [(1, 1024), (7, 512), (3, 37), (256, 1024)])
def test_quantizer_matches_oracle(cic, batch, L):
    rng = np.random.default_rng(L)
    lat = (rng.standard_normal((batch, L)) * 2).astype(np.float32)
    lat[0, : min(L, 8)] = [0.5, 1.5, 2.5, -0.5, -1.5, 3.5, 0.0, -2.5][: min(L, 8)]
    sal = rng.random((batch, 1)).astype(np.float32)
    sal[0] = 1.0                                                               # scale exactly 1 -> exact .5 ties
    qs = (0.1 + 0.8 * rng.random((batch, 1))).astype(np.float32)
    r = cic.ops.quantize_latent(lat, sal, qs, want=("deq", "symbols", "pre", "scale"))
    deq, sym, pre, scale = graphs.adaptive_quantize(lat, sal, qs)
    # exp() may differ by an ulp between libm and CUDA; symbols must agree except within 1e-3 of a tie
    got_sym = r["symbols"].cpu().numpy()
    bad = got_sym != sym.astype(np.int32)
    near_tie = np.abs(np.abs(pre - np.floor(pre)) - 0.5) < 1e-3
    assert not np.any(bad & ~near_tie), f"{int((bad & ~near_tie).sum())} symbol mismatches away from rounding boundaries"
    np.testing.assert_array_equal(got_sym[0], sym[0].astype(np.int32))          # exact-tie row: half-to-even
    np.testing.assert_allclose(r["scale"].cpu().numpy(), scale.ravel(), rtol=3e-7)
    ok = ~bad
    np.testing.assert_allclose(r["deq"].cpu().numpy()[ok], deq[ok], rtol=1e-6, atol=1e-7)


def test_quantizer_empty_batch(cic):
    r = cic.ops.quantize_latent(np.zeros((0, 16), np.float32), np.zeros((0,), np.float32), np.zeros((0,), np.float32))
    assert r["deq"].shape == (0, 16)


@pytest.mark.parametrize("h,w", [(256, 256), (64, 48), (17, 13)])
def test_roi_blend_matches_oracle(cic, h, w):
    rng = np.random.default_rng(h * w)
    n = 3
    mask = cic.synth.synth_masks(n, h, w) if min(h, w) >= 32 else rng.random((n, h, w, 1)).astype(np.float32)
    mask[0, 0, 0, 0] = 0.0
    hq = (rng.random((n, h, w, 3)) * 2 - 1).astype(np.float32)
    lq = (rng.random((n, h, w, 3)) * 2 - 1).astype(np.float32)
    bpp = np.array([0.1, 1.0, 2.0], np.float32)
    out, dt, s = cic.ops.roi_mask_blend(hq, lq, mask, bpp)
    odt = graphs.dynamic_threshold(mask, bpp)
    np.testing.assert_allclose(dt.cpu().numpy(), odt, atol=3e-6)
    np.testing.assert_allclose(out.cpu().numpy(), hq * odt + lq * (1.0 - odt), atol=5e-6)
    np.testing.assert_allclose(s.cpu().numpy() / (h * w), odt.reshape(n, -1).mean(1, dtype=np.float64), atol=1e-6)
    assert dt.cpu().numpy()[0, 0, 0, 0] < 1e-6                                   # pow(0, 0.7) = 0 path
    for row_bpp in ("0.1", "1.0", "2.0"):
        m = np.array(KA["dynamic_threshold"]["mask"], np.float32).reshape(1, -1, 1, 1)
        _, d, _ = cic.ops.roi_mask_blend(None, None, m, [float(row_bpp)])
        np.testing.assert_allclose(d.cpu().numpy().ravel(), KA["dynamic_threshold"][row_bpp], atol=2e-6)


def test_hq_ratio_sweep_matches_full_model_definition(cic):
    masks = cic.synth.synth_masks(5, 256, 256)
    levels = cic.synth.rate_control_bpps().astype(np.float32)
    got = cic.ops.hq_ratio_sweep(masks, levels).cpu().numpy()
    want = np.stack([graphs.dynamic_threshold(masks, np.full(5, b, np.float32)).reshape(5, -1).mean(1, dtype=np.float64)
                     for b in levels], 1)
    np.testing.assert_allclose(got, want, atol=1e-5)
    assert np.all(np.diff(got, axis=1) > 0)                                      # monotone like hq_ratio_by_bpp.png
    # odd pixel count exercises the scalar tail
    m2 = np.random.default_rng(0).random((2, 9, 7, 1)).astype(np.float32)
    got2 = cic.ops.hq_ratio_sweep(m2, levels[:3]).cpu().numpy()
    want2 = np.stack([graphs.dynamic_threshold(m2, np.full(2, b, np.float32)).reshape(2, -1).mean(1, dtype=np.float64)
                      for b in levels[:3]], 1)
    np.testing.assert_allclose(got2, want2, atol=1e-6)


def test_symbol_entropy(cic):
    rng = np.random.default_rng(4)
    sym = np.rint(rng.standard_normal((6, 1024)) * 5).astype(np.int32)
    sym[0] = 0
    sym[1, :512] = 3
    sym[1, 512:] = -3
    got = cic.ops.symbol_entropy_bits(torch.from_numpy(sym).cuda()).cpu().numpy()
    want = [metrics.symbol_entropy_bits(r) for r in sym]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9)
    assert got[0] == 0.0 and got[1] == pytest.approx(1024.0)


def test_u8_truncating_cast(cic):
    x = np.array([0.0, 0.999999, 1.0, 0.5, 127.9999 / 255, 128.0 / 255, 0.0039, 0.00393], np.float32)
    rng = np.random.default_rng(5)
    x = np.concatenate([x, rng.random(1001).astype(np.float32)])
    got = cic.ops.f32_to_u8_trunc(x, 255.0).cpu().numpy()
    np.testing.assert_array_equal(got, (x * np.float32(255)).astype(np.uint8))   # truncation, not rounding


@pytest.mark.parametrize("h,w", [(256, 256), (64, 80), (40, 33), (7, 7)])
def test_metrics_f32_match_oracle(cic, h, w):
    rng = np.random.default_rng(h + w)
    n = 3
    a = (cic.synth.to_signed_range(cic.synth.synth_images_u8(n, h, w, seed=9))).astype(np.float32)
    b = np.clip(a + rng.standard_normal(a.shape).astype(np.float32) * np.float32(0.08), -1, 1).astype(np.float32)
    got = cic.ops.metrics_f32(a, b, signed_range=True).cpu().numpy()
    for i in range(n):
        m = metrics.compute_metrics(a[i], b[i])
        assert abs(got[i, 0] - m["psnr"]) < 1e-6                                 # dB (north star: 0.05 dB)
        assert abs(got[i, 1] - m["ssim"]) < 2e-6                                 # stated tolerance 1e-4
        assert abs(got[i, 2] - float(m["mse"])) < 1e-8
    d = cic.gan.compute_metrics(a[0], b[0])
    assert set(d) == {"psnr", "ssim", "mse"} and abs(d["psnr"] - got[0, 0]) < 1e-9


@pytest.mark.parametrize("n,h,w", [(4, 416, 400), (7, 300, 333)])
def test_metrics_f32_many_tiles_match_oracle(cic, n, h, w):
    # several waves of tiles per SM, ragged right / bottom tiles
    rng = np.random.default_rng(n * h + w)
    a = (cic.synth.to_signed_range(cic.synth.synth_images_u8(n, h, w, seed=13))).astype(np.float32)
    b = np.clip(a + rng.standard_normal(a.shape).astype(np.float32) * np.float32(0.05), -1, 1).astype(np.float32)
    got = cic.ops.metrics_f32(a, b, signed_range=True).cpu().numpy()
    for i in range(n):
        m = metrics.compute_metrics(a[i], b[i])
        assert abs(got[i, 0] - m["psnr"]) < 1e-6
        assert abs(got[i, 1] - m["ssim"]) < 2e-6
        assert abs(got[i, 2] - float(m["mse"])) < 1e-8


@pytest.mark.parametrize("n,h,w", [(3, 256, 256), (2, 64, 80), (2, 40, 33), (1, 7, 7), (4, 416, 400)])
def test_metrics_f32_fast_ssim_match_oracle(cic, n, h, w):
    """cic_metrics_psnr_ssim_f32_fast: float32 window sums on centred data.  psnr / mse as the exact call; ssim within 1e-5 of
    the scipy / scikit-image arithmetic (survey criterion 1e-4)."""
    rng = np.random.default_rng(h * 7 + w)
    a = (cic.synth.to_signed_range(cic.synth.synth_images_u8(n, h, w, seed=17))).astype(np.float32)
    b = np.clip(a + rng.standard_normal(a.shape).astype(np.float32) * np.float32(0.06), -1, 1).astype(np.float32)
    got = cic.ops.metrics_f32(a, b, signed_range=True, fast=True).cpu().numpy()
    exact = cic.ops.metrics_f32(a, b, signed_range=True).cpu().numpy()
    np.testing.assert_allclose(got[:, [0, 2, 3]], exact[:, [0, 2, 3]], rtol=1e-12)   # psnr, mse, sse: same double sums (atomic order)
    for i in range(n):
        m = metrics.compute_metrics(a[i], b[i])
        assert abs(got[i, 1] - m["ssim"]) < 1e-5, (got[i, 1], m["ssim"])


def test_metrics_f32_fast_ssim_flat_bright_regions(cic):
    """Worst case for E[x^2] - E[x]^2: nearly constant bright images (variances ~1e-6 against means ~0.95) and a reconstruction
    that differs by a tiny structured error - the case the double accumulation of scipy exists for."""
    rng = np.random.default_rng(3)
    h = w = 128
    base = np.full((3, h, w, 3), 0.9, np.float32) + rng.standard_normal((3, h, w, 3)).astype(np.float32) * np.float32(1e-3)
    base[1, :, : w // 2] -= 1.7                                                   # a dark half: a step edge inside tiles
    base[2] = np.float32(0.97)                                                    # exactly constant
    a = np.clip(base, -1, 1).astype(np.float32)
    b = np.clip(a + rng.standard_normal(a.shape).astype(np.float32) * np.float32(2e-3), -1, 1).astype(np.float32)
    got = cic.ops.metrics_f32(a, b, signed_range=True, fast=True).cpu().numpy()
    for i in range(3):
        m = metrics.compute_metrics(a[i], b[i])
        assert abs(got[i, 1] - m["ssim"]) < 1e-5, (i, got[i, 1], m["ssim"])


def test_metric_sums_match_torch(cic):
    rng = np.random.default_rng(21)
    n = 37
    m = torch.from_numpy(rng.random((n, 4))).cuda()
    dts = torch.from_numpy(rng.random(n) * 512 * 512).cuda()
    got = cic.ops.metric_sums(m, dts, 512 * 512, 1024, 512, 256 * 256).cpu().numpy()[0]
    hq = dts.cpu().numpy() / (512 * 512)
    bpp = (hq * 1024 * 32 + (1 - hq) * 512 * 32) / 65536                          # GAN_test.py:313-318
    want = [m[:, 0].sum().item(), m[:, 1].sum().item(), m[:, 2].sum().item(), bpp.sum(), hq.sum(), 0.0, float(n), 0.0]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    assert abs(bpp.mean() - 0.25 * (1 + hq.mean())) < 1e-12                       # actual_bpp == 0.25 (1 + hq_ratio), SURVEY 8 a13


def test_metrics_identical_images(cic):
    a = cic.synth.to_signed_range(cic.synth.synth_images_u8(1, 32, 32))
    got = cic.ops.metrics_f32(a, a, signed_range=True).cpu().numpy()[0]
    assert np.isinf(got[0]) and got[1] == pytest.approx(1.0, abs=1e-7) and got[2] == 0.0


@pytest.mark.parametrize("h,w", [(256, 256), (128, 128), (50, 37)])
def test_metrics_gray_u8_match_oracle(cic, h, w):
    rng = np.random.default_rng(h)
    a = cic.synth.synth_images_u8(3, h, w, seed=11)
    b = np.clip(a.astype(int) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
    got = cic.ops.metrics_gray_u8(a, b).cpu().numpy()
    for i in range(3):
        assert abs(got[i, 0] - metrics.ae_calculate_psnr(a[i], b[i])) < 1e-9
        assert abs(got[i, 1] - metrics.ae_calculate_ssim(a[i], b[i])) < 1e-9     # float64 path: exact to rounding
        assert abs(got[i, 2] - metrics.ae_true_mse(a[i], b[i])) < 1e-9
        assert abs(got[i, 3] - metrics.ae_calculate_mse(a[i], b[i])) < 1e-9      # the reference's wrapped uint8 mse
    import test_autoencoder as ta
    assert abs(ta.calculate_ssim(a[0], b[0]) - got[0, 1]) < 1e-12
    assert abs(ta.calculate_psnr(a[0], b[0]) - got[0, 0]) < 1e-12
    assert abs(ta.calculate_mse(a[0], b[0]) - got[0, 3]) < 1e-12


CONV_CASES = [  # kh, stride, H, W, Cin, Cout, act
    (3, 1, 16, 24, 3, 32, "relu"), (3, 1, 16, 16, 32, 64, "relu"), (3, 1, 12, 20, 128, 32, "relu"),
    (4, 2, 32, 32, 3, 64, "lrelu"), (4, 2, 16, 16, 64, 128, "lrelu"), (4, 1, 16, 16, 32, 3, "tanh"),
    (1, 1, 8, 8, 256, 320, None), (3, 2, 32, 32, 1, 32, "lrelu"), (3, 2, 16, 16, 32, 64, "lrelu"),
    (3, 2, 9, 7, 16, 24, "sigmoid"), (3, 1, 5, 6, 5, 7, None),
]


@pytest.mark.parametrize("kh,stride,H,W,Cin,Cout,act", CONV_CASES)
def test_conv2d_same_matches_oracle(cic, kh, stride, H, W, Cin, Cout, act):
    rng = np.random.default_rng(kh * 100 + Cin)
    x = rng.standard_normal((2, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((kh, kh, Cin, Cout)) / np.sqrt(kh * kh * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * np.float32(0.1)
    scale = (0.5 + rng.random(Cout)).astype(np.float32)
    shift = (rng.random(Cout) - 0.5).astype(np.float32)
    got = cic.ops.conv2d_same(x, k, b, stride=stride, activation=act, scale=scale, shift=shift).cpu().numpy()
    t = graphs.conv2d_same(graphs._nchw(torch.from_numpy(x)), k, b, stride, torch.float32)
    t = t * torch.from_numpy(scale).view(1, -1, 1, 1) + torch.from_numpy(shift).view(1, -1, 1, 1)
    t = {"relu": torch.relu, "lrelu": graphs.lrelu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, None: lambda v: v}[act](t)
    np.testing.assert_allclose(got, graphs._nhwc(t).numpy(), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("H,W,Cin,Cout", [(4, 4, 512, 256), (8, 6, 64, 32), (5, 7, 16, 24)])
def test_conv_transpose_matches_oracle(cic, H, W, Cin, Cout):
    rng = np.random.default_rng(H * Cin)
    x = rng.standard_normal((2, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((4, 4, Cout, Cin)) / np.sqrt(4 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * np.float32(0.1)
    got = cic.ops.conv2d_transpose_k4s2(x, k, b, activation="lrelu").cpu().numpy()
    t = graphs.lrelu(graphs.conv2d_transpose_same_k4s2(graphs._nchw(torch.from_numpy(x)), k, b, torch.float32))
    np.testing.assert_allclose(got, graphs._nhwc(t).numpy(), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("B,K,N", [(1, 8192, 64), (5, 131072, 32), (256, 1024, 512), (3, 65, 128), (4, 256, 1), (2, 128, 3)])
def test_dense_matches_oracle(cic, B, K, N):
    rng = np.random.default_rng(K + N)
    x = rng.standard_normal((B, K)).astype(np.float32)
    k = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    got = cic.ops.dense(x, k, b, activation="relu").cpu().numpy()
    want = np.maximum(x.astype(np.float64) @ k.astype(np.float64) + b, 0)
    np.testing.assert_allclose(got, want, atol=3e-5 * np.sqrt(K / 64), rtol=1e-5)
    again = cic.ops.dense(x, k, b, activation="relu").cpu().numpy()
    np.testing.assert_array_equal(got, again)                                   # fixed-order split-K: bit-reproducible


@pytest.mark.parametrize("hw,C", [(8, 256), (32, 256), (4, 64)])
def test_self_attention_matches_oracle(cic, hw, C):
    rng = np.random.default_rng(hw)
    x = (rng.standard_normal((2, hw, hw, C)) * 0.5).astype(np.float32)
    w = {f"attn/{n}/kernel": (rng.standard_normal((1, 1, C, co)) / np.sqrt(C)).astype(np.float32)
         for n, co in (("query", C // 8), ("key", C // 8), ("value", C))}
    for n, co in (("query", C // 8), ("key", C // 8), ("value", C)):
        w[f"attn/{n}/bias"] = (rng.standard_normal(co) * 0.1).astype(np.float32)
    w["attn/gamma"] = np.array([0.5], np.float32)
    got = cic.ops.self_attention(x, w["attn/query/kernel"], w["attn/query/bias"], w["attn/key/kernel"], w["attn/key/bias"],
                                 w["attn/value/kernel"], w["attn/value/bias"], 0.5).cpu().numpy()
    want = graphs._nhwc(graphs.self_attention(w, graphs._nchw(torch.from_numpy(x)), torch.float32)).numpy()
    np.testing.assert_allclose(got, want, atol=3e-5, rtol=1e-5)
    ident = cic.ops.self_attention(x, w["attn/query/kernel"], None, w["attn/key/kernel"], None, w["attn/value/kernel"], None, 0.0)
    np.testing.assert_array_equal(ident.cpu().numpy(), x)                        # gamma = 0 -> exact identity


@pytest.mark.parametrize("n,h,w", [(2, 176, 208), (1, 256, 256), (1, 360, 500)])
def test_ms_ssim_matches_numpy_restatement(cic, n, h, w):
    """cic_msssim_f32 (BASELINE configs[4] extra, parity unpinned: no reference call site) against the float64 numpy restatement
    of Wang 2003 / tf.image.ssim_multiscale in oracle/metrics.py."""
    a = cic.synth.to_signed_range(cic.synth.synth_images_u8(n, h, w, seed=90))
    rng = np.random.default_rng(1)
    b = np.clip(a + rng.normal(0, 0.08, a.shape).astype(np.float32), -1, 1).astype(np.float32)
    got = cic.ops.ms_ssim_f32(a, b, signed_range=True).cpu().numpy()
    for i in range(n):
        want = metrics.ms_ssim((a[i] + 1) / 2, (b[i] + 1) / 2, data_range=1.0)
        assert abs(got[i] - want) < 1e-5, (got[i], want)
    same = cic.ops.ms_ssim_f32(a, a, signed_range=True).cpu().numpy()
    assert np.all(np.abs(same - 1.0) < 1e-6)
    with pytest.raises(ValueError, match="176"):
        cic.ops.ms_ssim_f32(a[:, :100], b[:, :100], signed_range=True)


@pytest.mark.parametrize("rows,L,scale", [(5, 1024, 3.0), (3, 70, 20.0), (2, 512, 0.0), (1, 32, 400.0), (9, 64, 1.0), (130, 512, 2.0)])
def test_rans_bitstream_matches_cpu_restatement(cic, rows, L, scale):
    """cic_rans_encode / cic_rans_decode: the GPU stream equals the numpy restatement byte for byte, both decoders
    invert both encoders, symbols beyond +-1023 are clamped.  Edge cases: one-symbol alphabet (scale 0), L not a multiple of 32,
    L < 32 lanes' worth, heavy tails."""
    from oracle import rans
    rng = np.random.default_rng(rows * 1000 + L)
    x = np.rint(rng.standard_normal((rows, L)) * scale).astype(np.int32)
    if scale > 100:
        x[0, :4] = [5000, -5000, 1023, -1023]
    want = rans.encode(x)
    got = cic.ops.rans_encode(torch.from_numpy(x).cuda())
    got_b = got.cpu().numpy().tobytes()
    assert len(got_b) == len(want)
    assert got_b == want
    clamped = np.clip(x, -1023, 1023)
    np.testing.assert_array_equal(cic.ops.rans_decode(got, rows, L).cpu().numpy(), clamped)
    np.testing.assert_array_equal(rans.decode(got_b), clamped)
    np.testing.assert_array_equal(cic.ops.rans_decode(torch.from_numpy(np.frombuffer(want, np.uint8).copy()), rows, L).cpu().numpy(), clamped)


def test_rans_round_trip_on_codec_symbols_and_empty(cic):
    """Round trip at the codec's own scale: symbols of 256 tiles (Laplacian-like, |s| <~ 100) -> stream -> symbols; measured bits
    per symbol sit between the zeroth-order entropy and entropy + 1.5 (32 x 32-bit states per row = 1 bit per symbol at L = 1024)."""
    rng = np.random.default_rng(7)
    x = np.rint(rng.laplace(0, 2.5, (256, 1024))).astype(np.int32)
    sym = torch.from_numpy(x).cuda()
    stream = cic.ops.rans_encode(sym)
    back = cic.ops.rans_decode(stream, 256, 1024)
    assert torch.equal(back, sym)
    bits = cic.ops.symbol_entropy_bits(sym.reshape(1, -1)).item()                 # zeroth-order entropy of the whole call
    coded = 8.0 * stream.numel()
    assert bits <= coded <= bits + 1.5 * x.size + 8 * 6000, (bits, coded)
    with pytest.raises(ValueError, match="CICR"):
        cic.ops.rans_decode(stream, 128, 1024)                                   # wrong shape for this stream
    with pytest.raises(ValueError, match="shorter"):
        cic.ops.rans_decode(stream[:100], 256, 1024)
    empty = cic.ops.rans_encode(torch.zeros((0, 64), dtype=torch.int32, device="cuda"))
    assert empty.numel() == 32 + 4096 + 4
    assert cic.ops.rans_decode(empty, 0, 64).shape == (0, 64)


@pytest.mark.parametrize("b,h,w", [(2, 256, 256), (1, 100, 180), (3, 33, 47)])
def test_saliency_mask_smooth_matches_opencv(cic, b, h, w):
    """cic_saliency_mask_smooth against the REAL OpenCV calls of create_saliency_mask(smooth=True)
    (GAN_functions.py:199-203): cv2.bilateralFilter(9, 75, 75) -> cv2.GaussianBlur((31, 31), 0) -> / max."""
    import cv2
    rng = np.random.default_rng(h + w)
    yy, xx = np.mgrid[0:h, 0:w]
    maps = []
    for _ in range(b):
        m = 0.05 * rng.random((h, w))
        for _ in range(4):
            cy, cx, s = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(4, 40)
            m += rng.uniform(0.2, 1.0) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
        maps.append((m / m.max()).astype(np.float32))
    maps = np.stack(maps)
    got = cic.ops.saliency_mask_smooth(maps).cpu().numpy()
    from oracle import saliency as osal
    np.testing.assert_array_equal(cic.saliency.create_saliency_mask(maps[0], smooth=True), got[0])   # the drop-in name: numpy in / out
    for i in range(b):
        want = osal.create_saliency_mask(maps[i], smooth=True)                    # the cv2 calls
        assert got[i].max() == pytest.approx(1.0, abs=1e-6)
        np.testing.assert_allclose(got[i], want, atol=1e-5)
    flat = np.full((1, 40, 40), 0.25, np.float32)                               # value range < FLT_EPSILON: bilateral copies, max-normalised to 1
    np.testing.assert_allclose(cic.ops.saliency_mask_smooth(flat).cpu().numpy(), 1.0, atol=1e-6)
    zero = np.zeros((40, 40), np.float32)                                       # max == 0: left as is (GAN_functions.py:202)
    assert cic.ops.saliency_mask_smooth(zero).abs().max().item() == 0.0


@pytest.mark.parametrize("b,h,w", [(3, 256, 256), (1, 100, 173), (2, 512, 384), (1, 1080, 1920), (2, 64, 64), (1, 37, 300)])
def test_saliency_map_matches_oracle(cic, b, h, w):
    """cic_saliency_map_u8 (compute_saliency_map, GAN_functions.py:52-121) against oracle/saliency.py: the numpy
    restatement (fine grained: the uint8 conspicuity map bit for bit; spectral residual: float64 pipeline, 1e-5) and the same
    detectors composed of the REAL OpenCV core routines (1e-3: cv2's polar conversions are float32)."""
    from oracle import saliency as osal
    from test_oracle_saliency import photo
    bgr = np.stack([photo(h, w, seed=10 * h + i) for i in range(b)])
    rgb_u8 = np.ascontiguousarray(bgr[..., ::-1])
    signed = (rgb_u8.astype(np.float32) - 127.5) / 127.5                         # what load_and_preprocess_image returns
    for method in ("spectral_residual", "fine_grained", "combined"):
        got = cic.ops.saliency_map(signed, method).cpu().numpy()
        got_u8 = cic.ops.saliency_map(rgb_u8, method).cpu().numpy()              # uint8 input: cast as is (:65-66)
        assert got.shape == (b, h, w) and got.dtype == np.float32
        for i in range(b):
            want = osal.compute_saliency_map(signed[i], method)
            tol = 0 if method == "fine_grained" else 1e-5
            np.testing.assert_allclose(got[i], want, atol=tol, rtol=0)
            np.testing.assert_allclose(got[i], osal.compute_saliency_map(signed[i], method, use_cv=True), atol=1e-3)
            assert got[i].max() == pytest.approx(1.0, abs=1e-6)
            # ((u8 - 127.5) / 127.5 + 1) * 127.5 truncates back to u8 or u8 - 1: the two input conventions agree closely, not exactly
            assert np.mean(np.abs(got_u8[i] - got[i])) < 2e-2
    single = cic.saliency.compute_saliency_map(signed[0], method="combined")     # the drop-in name: numpy in, numpy out
    np.testing.assert_array_equal(single, got[0])
    with pytest.raises(ValueError, match="Unsupported"):
        cic.ops.saliency_map(signed, "nope")


def test_saliency_map_degenerate_inputs(cic):
    """A flat image: the fine-grained map is all zero and stays zero (no division by a zero maximum, GAN_functions.py:118); the
    whole front end on a batch equals per-image calls."""
    from oracle import saliency as osal
    from test_oracle_saliency import photo
    flat = np.full((1, 48, 48, 3), 90, np.uint8)
    assert cic.ops.saliency_map(flat, "fine_grained").abs().max().item() == 0.0
    bgr = np.stack([photo(128, 160, seed=i) for i in range(3)])
    rgb = np.ascontiguousarray(bgr[..., ::-1])
    masks = cic.ops.saliency_mask_from_image(rgb).cpu().numpy()
    for i in range(3):
        one = cic.ops.saliency_mask_from_image(rgb[i]).cpu().numpy()
        np.testing.assert_array_equal(masks[i], one)
        want = osal.create_saliency_mask(osal.compute_saliency_map(rgb[i], "combined"), smooth=True)
        np.testing.assert_allclose(masks[i], want, atol=1e-4)


def test_saliency_mask_binary_matches_opencv_and_numpy(cic):
    """cic_saliency_mask_binary (create_saliency_mask(smooth=False), GAN_functions.py:172-197, :204-206) against the reference's own
    lines run with the REAL cv2.threshold(THRESH_OTSU) and np.histogram (oracle.saliency.adaptive_threshold)."""
    from oracle import saliency as osal
    rng = np.random.default_rng(5)
    maps = []
    for t in range(12):
        h, w = 40 + 13 * t, 64 + 7 * t
        kind = t % 4
        if kind == 0:
            m = rng.random((h, w)) ** rng.uniform(0.3, 4)
        elif kind == 1:
            m = (rng.random((h, w)) > rng.random()) * rng.random()
        elif kind == 2:
            m = np.clip(rng.normal(rng.random(), 0.2, (h, w)), 0, 1)
        else:
            m = np.round(rng.random((h, w)) * 8) / 8              # values on bin edges (0.125 k), 1.0 included
        maps.append(m.astype(np.float32))
    for m in maps:
        want_thr = osal.adaptive_threshold(m)
        got, thr = cic.ops.saliency_mask_binary(m, None, return_threshold=True)
        assert thr[0].item() == want_thr
        np.testing.assert_array_equal(got.cpu().numpy(), osal.create_saliency_mask(m, smooth=False))
        np.testing.assert_array_equal(cic.saliency.create_saliency_mask(m, smooth=False), got.cpu().numpy())
        assert cic.saliency.adaptive_threshold(m) == want_thr
        np.testing.assert_array_equal(cic.saliency.create_saliency_mask(m, threshold=0.3, smooth=False), (m > 0.3).astype(np.float32))
    batch = np.stack([rng.random((48, 80)).astype(np.float32) ** (1 + i) for i in range(5)])
    got, thr = cic.ops.saliency_mask_binary(batch, None, return_threshold=True)
    for i in range(5):
        assert thr[i].item() == osal.adaptive_threshold(batch[i])
        np.testing.assert_array_equal(got[i].cpu().numpy(), (batch[i] > thr[i].item()).astype(np.float32))
    big = (rng.random((32, 32)) * 300).astype(np.float32)         # maximum above 1: cast to uint8 as is (:176-177), wraps like numpy
    assert cic.saliency.adaptive_threshold(big) == osal.adaptive_threshold(big)


@pytest.mark.parametrize("b,h,w", [(2, 128, 160), (1, 37, 61)])
def test_saliency_enhance_matches_opencv(cic, b, h, w):
    """cic_saliency_enhance (enhance_saliency_map, GAN_functions.py:123-157) against the reference's lines on the real OpenCV."""
    from oracle import saliency as osal
    rng = np.random.default_rng(h)
    yy, xx = np.mgrid[0:h, 0:w]
    maps = np.stack([(0.05 * rng.random((h, w)) + np.exp(-((yy - rng.uniform(0, h)) ** 2 + (xx - rng.uniform(0, w)) ** 2) / (2 * 20.0 ** 2))).astype(np.float32)
                     for _ in range(b)])
    maps /= maps.max()
    got = cic.ops.saliency_enhance(maps).cpu().numpy()
    for i in range(b):
        np.testing.assert_allclose(got[i], osal.enhance_saliency_map(maps[i]), atol=2e-6)
    import GAN_functions as gf
    np.testing.assert_array_equal(gf.enhance_saliency_map(maps[0]), got[0])


def _jpeg_image(h, w, kind, seed):
    from test_oracle_extras import _jpeg_test_image
    return _jpeg_test_image(h, w, kind, seed)


@pytest.mark.parametrize("h,w", [(16, 16), (64, 64), (100, 150), (17, 33), (8, 8), (1, 1), (120, 68), (256, 256), (512, 512)])
def test_jpeg_encoder_is_byte_identical_to_opencv(cic, h, w):
    """cic_jpeg_encode_u8 (the output stage) against the REAL library behind the reference's cv2.imwrite
    (test_autoencoder.py:93; GAN_functions.py:50): the files are equal byte for byte - and equal to the numpy restatement."""
    import cv2
    from oracle import jpeg
    imgs = np.stack([_jpeg_image(h, w, kind, h * 1000 + w + i) for i, kind in enumerate(("noise", "smooth", "flat", "saturated", "smooth"))])
    files = cic.ops.jpeg_encode(imgs)
    assert len(files) == len(imgs)
    for i, f in enumerate(files):
        want = bytes(cv2.imencode(".jpg", imgs[i])[1])
        assert f[:jpeg.HEADER_BYTES] == want[:jpeg.HEADER_BYTES], "header"
        assert len(f) == len(want), (i, len(f), len(want))
        assert f == want, f"image {i}: first difference at byte {next(k for k in range(len(f)) if f[k] != want[k])} of {len(f)}"
    assert files[1] == jpeg.encode_bgr(imgs[1])
    for q in (100, 50, 10):
        assert cic.ops.jpeg_encode(imgs[1], quality=q) == bytes(cv2.imencode(".jpg", imgs[1], [cv2.IMWRITE_JPEG_QUALITY, q])[1]), q
    assert cic.ops.jpeg_encode(imgs[0], quality=100) == bytes(cv2.imencode(".jpg", imgs[0], [cv2.IMWRITE_JPEG_QUALITY, 100])[1])   # longest codes
    # RGB input = save_image's cv2.cvtColor(img, cv2.COLOR_RGB2BGR) folded into the kernel
    assert cic.ops.jpeg_encode(imgs[1][..., ::-1].copy(), rgb=True) == files[1]


def test_jpeg_encoder_frames_and_capacity(cic):
    import cv2
    frame = _jpeg_image(1080, 1920, "smooth", 3)                                 # H is not a multiple of 16: dummy block row
    f = cic.ops.jpeg_encode(torch.from_numpy(frame).cuda())
    assert f == bytes(cv2.imencode(".jpg", frame)[1])
    back = cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR)
    assert back.shape == frame.shape
    # a buffer that is too small: the size is still reported, nothing is written beyond the capacity
    x = torch.from_numpy(np.stack([_jpeg_image(64, 64, "noise", 1)] * 2)).cuda()
    full, sizes = cic.ops.jpeg_encode_device(x)
    need = int(sizes[0].item())
    out, sizes2 = cic.ops.jpeg_encode_device(x, capacity=1000)
    assert out.shape == (2, 1000) and int(sizes2[0].item()) == need > 1000
    assert torch.equal(out[0], full[0, :1000]) and torch.equal(out[1], full[1, :1000])
    with pytest.raises(ValueError):
        cic.ops.jpeg_encode_device(x.float())
    empty, s0 = cic.ops.jpeg_encode_device(x[:0])
    assert empty.shape[0] == 0 and s0.numel() == 0


def test_save_image_writes_opencv_bytes_from_the_gpu(cic, tmp_path):
    """GAN_functions.save_image(img, path) (:41-50): [-1,1] RGB float -> truncated uint8 -> BGR -> cv2.imwrite.  For .jpg paths the file
    is made on the GPU and equals the one OpenCV writes."""
    import cv2
    import GAN_functions as gf
    img = _jpeg_image(96, 80, "smooth", 4).astype(np.float32) / 127.5 - 1.0
    p = str(tmp_path / "a.jpg")
    gf.save_image(img, p)
    u8 = cv2.cvtColor(((img + 1) * 127.5).astype(np.uint8), cv2.COLOR_RGB2BGR)
    q = str(tmp_path / "b.jpg")
    cv2.imwrite(q, u8)
    assert open(p, "rb").read() == open(q, "rb").read()
    gf.save_image(img, str(tmp_path / "c.png"))                                  # other formats stay on OpenCV
    np.testing.assert_array_equal(cv2.imread(str(tmp_path / "c.png")), u8)


def test_autoencoder_batch_writes_the_files_opencv_would(cic, tmp_path):
    """test_autoencoder.py:88-93: (compressed * 255).astype(uint8) -> cv2.imwrite(compressed_path, ...) per image; evaluate_batch does
    it for the batch, .jpg on the GPU."""
    import cv2
    import train_autoencoder as tr
    model = tr.build_autoencoder((64, 96, 3))
    model.set_weights_dict(cic.weights.synthetic_autoencoder(seed=42))
    x = cic.synth.to_unit_range(cic.synth.synth_images_u8(3, 64, 96, seed=3))
    paths = [str(tmp_path / "a.jpg"), str(tmp_path / "b.png"), str(tmp_path / "c.jpeg")]
    r = cic.autoencoder.evaluate_batch(model, x, save_paths=paths)
    y8 = r["compressed_u8"].cpu().numpy()
    for i, p in enumerate(paths):
        q = str(tmp_path / ("ref_" + os.path.basename(p)))
        cv2.imwrite(q, y8[i])
        assert open(p, "rb").read() == open(q, "rb").read(), p
    with pytest.raises(ValueError):
        cic.autoencoder.evaluate_batch(model, x, save_paths=paths[:2])
