"""Check the torch graphs of the oracle against an independent numpy restatement of TF 'same' padding,
Keras kernel layouts and the transposed convolution, plus structural identities of the reference graphs."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import graphs

W = importlib.import_module("contextual-image-compression_b200.weights")
synth = importlib.import_module("contextual-image-compression_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))


def np_conv_same(x, k, b, stride):
    """Direct TF-'same' convolution on NHWC numpy (float64)."""
    B, H, Wd, Cin = x.shape
    kh, kw, _, Cout = k.shape
    Ho, Wo = -(-H // stride), -(-Wd // stride)
    pt = max((Ho - 1) * stride + kh - H, 0) // 2
    pl = max((Wo - 1) * stride + kw - Wd, 0) // 2
    y = np.zeros((B, Ho, Wo, Cout))
    for oy in range(Ho):
        for ox in range(Wo):
            for ky in range(kh):
                for kx in range(kw):
                    iy, ix = oy * stride + ky - pt, ox * stride + kx - pl
                    if 0 <= iy < H and 0 <= ix < Wd:
                        y[:, oy, ox, :] += x[:, iy, ix, :] @ k[ky, kx]
    return y + b


def np_deconv_k4s2(x, k, b):
    """Keras Conv2DTranspose(k4,s2,'same'): scatter form, output index o = 2i + k - 1; kernel (kh,kw,Cout,Cin)."""
    B, H, Wd, Cin = x.shape
    Cout = k.shape[2]
    y = np.zeros((B, 2 * H, 2 * Wd, Cout))
    for iy in range(H):
        for ix in range(Wd):
            for ky in range(4):
                for kx in range(4):
                    oy, ox = 2 * iy + ky - 1, 2 * ix + kx - 1
                    if 0 <= oy < 2 * H and 0 <= ox < 2 * Wd:
                        y[:, oy, ox, :] += x[:, iy, ix, :] @ k[ky, kx].T
    return y + b


@pytest.mark.parametrize("kh,stride,H,Wd", [(3, 1, 6, 7), (4, 2, 8, 6), (4, 1, 5, 6), (3, 2, 8, 8), (3, 2, 7, 9), (1, 1, 4, 4)])
def test_conv_same_padding(kh, stride, H, Wd):
    rng = np.random.default_rng(kh * 10 + stride)
    x = rng.standard_normal((2, H, Wd, 3))
    k = rng.standard_normal((kh, kh, 3, 4))
    b = rng.standard_normal(4)
    t = graphs._nchw(torch.from_numpy(x))
    y = graphs._nhwc(graphs.conv2d_same(t, k, b, stride, torch.float64)).numpy()
    np.testing.assert_allclose(y, np_conv_same(x, k, b, stride), atol=1e-12)


def test_conv_transpose_layout_and_crop():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3, 4, 5))
    k = rng.standard_normal((4, 4, 6, 5))
    b = rng.standard_normal(6)
    y = graphs._nhwc(graphs.conv2d_transpose_same_k4s2(graphs._nchw(torch.from_numpy(x)), k, b, torch.float64)).numpy()
    np.testing.assert_allclose(y, np_deconv_k4s2(x, k, b), atol=1e-12)


def test_attention_identity_at_gamma_zero_and_skip_pre_attention():
    img_shape = (32, 32, 3)
    w = W.synthetic_encoder(img_shape, 16, True, seed=1)
    img = synth.to_signed_range(synth.synth_images_u8(1, 32, 32))
    w0 = dict(w); w0["attn/gamma"] = np.zeros(1, np.float32)
    with_attn = graphs.encoder_forward(w, img, True)
    gamma0 = graphs.encoder_forward(w0, img, True)
    no_attn = graphs.encoder_forward(w, img, False)
    np.testing.assert_array_equal(gamma0[0], no_attn[0])         # gamma = 0 -> exact identity (App. D.2)
    assert np.abs(with_attn[0] - no_attn[0]).max() > 1e-4        # gamma = 0.5 changes the latent ...
    np.testing.assert_array_equal(with_attn[3], no_attn[3])      # ... but never skip3 (tapped before attention)


def test_graph_output_ranges_and_shapes():
    img_shape, base = (64, 64, 3), 32
    ws = W.synthetic_adaptive(img_shape, base, seed=7)
    img = synth.to_signed_range(synth.synth_images_u8(2, 64, 64))
    mask = synth.synth_masks(2, 64, 64)
    outs = graphs.adaptive_forward(ws, img, mask, np.array([[0.5], [1.5]], np.float32))
    assert [o.shape for o in outs] == [(2, 64, 64, 3), (2, 64), (2, 32), (2, 3), (2, 64, 64, 1)]
    assert np.abs(outs[0]).max() < 1 and 0 <= outs[4].min() and outs[4].max() <= 1
    assert np.all((outs[3] > 0) & (outs[3] < 1))
    ae = graphs.autoencoder_forward(W.synthetic_autoencoder(), synth.to_unit_range(synth.synth_images_u8(1, 16, 24)))
    assert ae.shape == (1, 16, 24, 3) and 0 < ae.min() and ae.max() < 1


def test_batchnorm_at_init_scale():
    x = torch.ones(1, 4, 2, 2)
    w = {"bn/gamma": np.ones(4, np.float32), "bn/beta": np.zeros(4, np.float32), "bn/moving_mean": np.zeros(4, np.float32),
         "bn/moving_variance": np.ones(4, np.float32)}
    assert graphs.batchnorm(x, w, "bn", torch.float32)[0, 0, 0, 0].item() == pytest.approx(0.999500394, abs=1e-7)


def test_default_init_quantises_everything_to_zero():
    """at Keras-default init every symbol is 0 - the reason for the synthetic recipe."""
    img_shape = (64, 64, 3)
    w = W.synthetic_encoder(img_shape, 32, False, seed=3, keras_default=True)
    lat = graphs.encoder_forward(w, synth.to_signed_range(synth.synth_images_u8(1, 64, 64)), False)[0]
    assert lat.std() < 0.05 and np.abs(lat * 3.03).max() < 0.5   # every symbol 0 at a typical scale
    w2 = W.synthetic_encoder(img_shape, 32, False, seed=3)
    lat2 = graphs.encoder_forward(w2, synth.to_signed_range(synth.synth_images_u8(1, 64, 64)), False)[0]
    assert 0.5 < lat2.std() < 6


def test_oracle_matches_frozen_golden():
    g = np.load(os.path.join(HERE, "golden", "oracle_small.npz"))
    torch.set_num_threads(1)
    y = graphs.autoencoder_forward(W.synthetic_autoencoder(seed=42), synth.to_unit_range(synth.synth_images_u8(2, 32, 48, seed=43)))
    np.testing.assert_allclose(y, g["ae_y"], atol=2e-6)
    ws = W.synthetic_adaptive((64, 64, 3), 32, seed=42)
    img = synth.to_signed_range(synth.synth_images_u8(3, 64, 64, seed=44))
    mask = synth.synth_masks(3, 64, 64, seed=44)
    outs, ex = graphs.adaptive_forward(ws, img, mask, np.array([[0.1], [1.0], [2.0]], np.float32), return_extras=True)
    np.testing.assert_allclose(outs[0], g["ad_blended"], atol=5e-5)
    np.testing.assert_allclose(ex["hq_latent"], g["ad_hq_latent"], atol=5e-5)
    np.testing.assert_allclose(outs[4], g["ad_dt"], atol=1e-6)
    assert (ex["hq_sym"] != g["ad_hq_sym"]).mean() < 0.02
