"""torch-CPU restatement of the reference's model graphs.  TEST INFRASTRUCTURE (see oracle/__init__).

All tensors at the interface are numpy, NHWC, like the Keras models of the reference.  `dtype`
is torch.float32 for the reference-equivalent result and torch.float64 for the ground truth
used to measure the fp32 noise floor and to classify symbol mismatches near rounding
boundaries.  Weights are dicts in Keras layouts (see the product package's weights.py).

Third-party semantics restated here: TF padding='same' (extra pad at the
end), Keras kernel layouts, BatchNormalization eps 1e-3 in inference mode, LeakyReLU 0.2,
UpSampling2D nearest, tf.round = half-to-even.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _nchw(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


def _nhwc(x_nchw: torch.Tensor) -> torch.Tensor:
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def same_pads(size: int, k: int, s: int):
    """TF 'same': out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, kernel, bias, stride, dtype):
    """Keras Conv2D(padding='same') on an NCHW tensor; kernel (kh,kw,Cin,Cout)."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    pt, pb = same_pads(x.shape[2], kh, stride)
    pl, pr = same_pads(x.shape[3], kw, stride)
    x = F.pad(x, (pl, pr, pt, pb))
    w = _t(kernel, dtype).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, w, _t(bias, dtype), stride=stride)


def conv2d_transpose_same_k4s2(x, kernel, bias, dtype):
    """Keras Conv2DTranspose(k=4, s=2, padding='same'); kernel (kh,kw,Cout,Cin).

    Output is 2x the input; TF crops 1 px each side of the full transposed convolution, which is
    ConvTranspose2d(padding=1) with weight[Cin,Cout,kh,kw] = kernel.transpose(3,2,0,1), no flip.
    """
    w = _t(kernel, dtype).permute(3, 2, 0, 1).contiguous()
    return F.conv_transpose2d(x, w, _t(bias, dtype), stride=2, padding=1)


def batchnorm(x, w, prefix, dtype):
    g = _t(w[prefix + "/gamma"], dtype).view(1, -1, 1, 1)
    b = _t(w[prefix + "/beta"], dtype).view(1, -1, 1, 1)
    m = _t(w[prefix + "/moving_mean"], dtype).view(1, -1, 1, 1)
    v = _t(w[prefix + "/moving_variance"], dtype).view(1, -1, 1, 1)
    return (x - m) / torch.sqrt(v + BN_EPS) * g + b


def lrelu(x):
    return F.leaky_relu(x, 0.2)


# --------------------------------------------------------------------------------------------
# autoencoder: train_autoencoder.py:9-40
# --------------------------------------------------------------------------------------------
def autoencoder_forward(w, x_nhwc: np.ndarray, dtype=torch.float32) -> np.ndarray:
    x = _nchw(_t(x_nhwc, dtype))
    c = lambda name, t: conv2d_same(t, w[name + "/kernel"], w[name + "/bias"], 1, dtype)
    x1 = F.relu(c("conv1", x))                                   # :14
    x1p = F.max_pool2d(x1, 2, ceil_mode=True)                     # :15  MaxPooling2D((2,2), 'same')
    x2 = F.relu(c("conv2", x1p))                                  # :17
    enc = F.max_pool2d(x2, 2, ceil_mode=True)                     # :18
    y = F.relu(c("conv3", enc))                                   # :21
    y = F.interpolate(y, scale_factor=2, mode="nearest")          # :22
    x2r = F.relu(c("conv_x2", x2))                                # :25
    y = torch.cat([y, x2r], dim=1)                                # :26
    y = F.relu(c("conv5", y))                                     # :28
    y = F.interpolate(y, scale_factor=2, mode="nearest")          # :29
    x1r = F.relu(c("conv_x1", x1))                                # :32
    y = torch.cat([y, x1r], dim=1)                                # :33
    y = torch.sigmoid(c("conv_out", y))                           # :35
    return _nhwc(y).numpy()


def autoencoder_output_u8(y: np.ndarray) -> np.ndarray:
    """`(compressed_img * 255).astype("uint8")` - truncation (test_autoencoder.py:88)."""
    return (y.astype(np.float32) * np.float32(255)).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# GAN encoder / attention / generator: GAN_functions.py:236-374
# --------------------------------------------------------------------------------------------
def self_attention(w, x, dtype):
    """SelfAttention.call (GAN_functions.py:344-369) on NCHW x with C=256; no 1/sqrt(d) scale."""
    b, c, h, wd = x.shape
    q = conv2d_same(x, w["attn/query/kernel"], w["attn/query/bias"], 1, dtype)
    k = conv2d_same(x, w["attn/key/kernel"], w["attn/key/bias"], 1, dtype)
    v = conv2d_same(x, w["attn/value/kernel"], w["attn/value/bias"], 1, dtype)
    q = _nhwc(q).reshape(b, h * wd, -1)
    k = _nhwc(k).reshape(b, h * wd, -1)
    v = _nhwc(v).reshape(b, h * wd, -1)
    amap = torch.softmax(torch.matmul(q, k.transpose(1, 2)), dim=-1)   # :358-359
    ctx = torch.matmul(amap, v).reshape(b, h, wd, c)                     # :363-364
    gamma = _t(w["attn/gamma"], dtype)
    return _nchw(gamma * ctx) + x                                        # :367


def encoder_forward(w, img_nhwc: np.ndarray, add_attention: bool, dtype=torch.float32):
    """build_encoder (GAN_functions.py:280-331) -> [latent, x1, x2, x3] (numpy, NHWC)."""
    x = _nchw(_t(img_nhwc, dtype))
    x1 = lrelu(conv2d_same(x, w["conv1/kernel"], w["conv1/bias"], 2, dtype))                  # :300-301
    x2 = lrelu(batchnorm(conv2d_same(x1, w["conv2/kernel"], w["conv2/bias"], 2, dtype), w, "bn2", dtype))
    x3 = lrelu(batchnorm(conv2d_same(x2, w["conv3/kernel"], w["conv3/bias"], 2, dtype), w, "bn3", dtype))
    x3a = self_attention(w, x3, dtype) if add_attention else x3                                 # skip is pre-attention (:312)
    x4 = lrelu(batchnorm(conv2d_same(x3a, w["conv4/kernel"], w["conv4/bias"], 2, dtype), w, "bn4", dtype))
    flat = _nhwc(x4).reshape(x4.shape[0], -1)                                                  # Flatten, NHWC (:325)
    latent = flat @ _t(w["dense/kernel"], dtype) + _t(w["dense/bias"], dtype)                  # :326
    return [latent.numpy(), _nhwc(x1).numpy(), _nhwc(x2).numpy(), _nhwc(x3).numpy()]


def generator_forward(w, latent, skip1, skip2, skip3, dtype=torch.float32) -> np.ndarray:
    """build_generator (GAN_functions.py:236-278); inputs numpy NHWC; output NHWC in (-1,1)."""
    z = _t(latent, dtype)
    h16, w16 = skip3.shape[1] // 2, skip3.shape[2] // 2
    x = z @ _t(w["dense/kernel"], dtype) + _t(w["dense/bias"], dtype)                          # :247
    x = _nchw(x.reshape(-1, h16, w16, 512))                                                    # :248 Reshape NHWC
    x = lrelu(batchnorm(x, w, "bn0", dtype))                                                   # :249-250
    for i, skip in ((1, skip3), (2, skip2), (3, skip1), (4, None)):
        x = conv2d_transpose_same_k4s2(x, w[f"deconv{i}/kernel"], w[f"deconv{i}/bias"], dtype)
        x = lrelu(batchnorm(x, w, f"bn{i}", dtype))
        if skip is not None:
            x = torch.cat([x, _nchw(_t(skip, dtype))], dim=1)                                  # :256,261,266
    y = torch.tanh(conv2d_same(x, w["conv_out/kernel"], w["conv_out/bias"], 1, dtype))          # :273
    return _nhwc(y).numpy()


def latent_saliency_forward(w, latent, dtype=torch.float32) -> np.ndarray:
    """build_latent_saliency_model (GAN_functions.py:210-234) -> (B,1)."""
    x = _t(latent, dtype)
    x = F.relu(x @ _t(w["dense1/kernel"], dtype) + _t(w["dense1/bias"], dtype))
    x = F.relu(x @ _t(w["dense2/kernel"], dtype) + _t(w["dense2/bias"], dtype))
    x = torch.sigmoid(x @ _t(w["dense3/kernel"], dtype) + _t(w["dense3/bias"], dtype))
    return x.numpy()


def rate_scalars(target_bpp, dtype=torch.float32):
    """t, hq_lq_threshold, quant_strength (GAN_functions.py:631-649), each (B,1)."""
    b = _t(np.asarray(target_bpp).reshape(-1, 1), dtype)
    t = torch.clamp(b / 5.0, 0.0, 1.0)
    return t, 0.9 - 0.85 * t, 0.9 - 0.8 * t


def adaptive_quantize(latent, saliency, quant_strength, dtype=torch.float32):
    """AdaptiveQuantizationLayer.call (GAN_functions.py:435-446).

    Returns (dequantised, symbols, pre_round, scale); symbols = round(latent*scale) half-to-even.
    """
    lat = _t(latent, dtype)
    sal = _t(saliency, dtype)
    qs = quant_strength if isinstance(quant_strength, torch.Tensor) else _t(quant_strength, dtype)
    eff = qs * (1.0 - sal)
    scale = torch.exp(eff * 3.0)
    pre = lat * scale
    sym = torch.round(pre)
    return (sym / scale).numpy(), sym.numpy(), pre.numpy(), scale.numpy()


def rd_optimizer_forward(w, mask_nhwc, target_bpp, dtype=torch.float32) -> np.ndarray:
    """build_rate_distortion_optimizer (GAN_functions.py:495-557) -> rd_params (B,3).

    The image input of the reference model is unused (:500), so it is not an argument here.
    """
    t, _, _ = rate_scalars(target_bpp, dtype)
    x = _nchw(_t(mask_nhwc, dtype))
    x = lrelu(conv2d_same(x, w["conv1/kernel"], w["conv1/bias"], 2, dtype))                    # :511-512
    x = lrelu(conv2d_same(x, w["conv2/kernel"], w["conv2/bias"], 2, dtype))                    # :513-514
    x = x.mean(dim=(2, 3))                                                                     # :515 GAP
    x = torch.cat([x, t], dim=1)                                                               # :518
    x = lrelu(x @ _t(w["dense1/kernel"], dtype) + _t(w["dense1/bias"], dtype))                 # :521-522
    p = x @ _t(w["dense2/kernel"], dtype) + _t(w["dense2/bias"], dtype)                        # :525
    out = torch.cat([torch.sigmoid(p[:, 0:1] + 1.0 - 2.0 * t),                                 # :529-531
                     torch.sigmoid(p[:, 1:2] + 1.0 - 2.0 * t),                                 # :534-536
                     torch.sigmoid(p[:, 2:3] + 1.0 - 1.5 * t)], dim=1)                         # :539-541
    return out.numpy()


def dynamic_threshold(mask_nhwc, target_bpp, dtype=torch.float32) -> np.ndarray:
    """dt = sigmoid((mask^0.7 - thr) * 20) (GAN_functions.py:651-657) -> (B,H,W,1)."""
    _, thr, _ = rate_scalars(target_bpp, dtype)
    m = _t(mask_nhwc, dtype)
    es = torch.pow(m, 0.7)
    return torch.sigmoid((es - thr.view(-1, 1, 1, 1)) * 20.0).numpy()


def adaptive_forward(weights, img_nhwc, mask_nhwc, target_bpp, dtype=torch.float32, return_extras=False):
    """The adaptive model of build_adaptive_compression_model (GAN_functions.py:604-696).

    Returns [blended, hq_latent_quantized, lq_latent_quantized, rd_params, dynamic_threshold]
    (+ a dict of intermediates when return_extras).
    """
    hq = encoder_forward(weights["hq_encoder"], img_nhwc, True, dtype)                          # :604
    lq = encoder_forward(weights["lq_encoder"], img_nhwc, False, dtype)                         # :605
    sal_hq = latent_saliency_forward(weights["latent_saliency_hq"], hq[0], dtype)               # :619
    sal_lq = latent_saliency_forward(weights["latent_saliency_lq"], lq[0], dtype)               # :620
    rd = rd_optimizer_forward(weights["rd_optimizer"], mask_nhwc, target_bpp, dtype)            # :624
    _, _, qs = rate_scalars(target_bpp, dtype)
    hq_q, hq_sym, hq_pre, hq_scale = adaptive_quantize(hq[0], sal_hq, qs, dtype)                # :665
    lq_q, lq_sym, lq_pre, lq_scale = adaptive_quantize(lq[0], sal_lq, qs, dtype)                # :666
    hq_out = generator_forward(weights["hq_generator"], hq_q, hq[1], hq[2], hq[3], dtype)       # :669
    lq_out = generator_forward(weights["lq_generator"], lq_q, lq[1], lq[2], lq[3], dtype)       # :670
    dt = dynamic_threshold(mask_nhwc, target_bpp, dtype)                                        # :655-657
    blended = hq_out * dt + lq_out * (1.0 - dt)                                                 # :682-684
    outs = [blended, hq_q, lq_q, rd, dt]
    if return_extras:
        extras = dict(hq_latent=hq[0], lq_latent=lq[0], hq_skips=hq[1:], lq_skips=lq[1:], sal_hq=sal_hq,
                      sal_lq=sal_lq, hq_sym=hq_sym, lq_sym=lq_sym, hq_pre=hq_pre, lq_pre=lq_pre,
                      hq_scale=hq_scale, lq_scale=lq_scale, hq_out=hq_out, lq_out=lq_out)
        return outs, extras
    return outs
