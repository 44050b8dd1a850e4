"""Weight containers in Keras layouts + the synthetic-weight recipe.

Layouts follow Keras so that the `.h5` importer (keras_h5.py) is a pure rename:
  Conv2D kernel           (kh, kw, Cin, Cout)      bias (Cout,)
  Conv2DTranspose kernel  (kh, kw, Cout, Cin)      bias (Cout,)
  Dense kernel            (in, out)                bias (out,)
  BatchNormalization      gamma, beta, moving_mean, moving_variance (C,), epsilon 1e-3
Flatten / Reshape are in NHWC order (`GAN_functions.py:248,325`).

The reference ships no checkpoints and Keras-default initialisation makes every quantised
symbol zero, so `synthetic_*` apply the documented recipe:
glorot-uniform kernels, small random biases, randomised BatchNorm statistics, the encoder's
final Dense scaled so the latent has std ~ 2, attention gamma = 0.5.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon

Weights = Dict[str, np.ndarray]
_SHAPES_ONLY = False


def _glorot(rng: np.random.Generator, shape: Tuple[int, ...]) -> np.ndarray:
    """Keras glorot_uniform: U(+-sqrt(6 / (fan_in + fan_out))) with fans from the kernel shape."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        receptive = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
    if _SHAPES_ONLY:   # check_adaptive wants the shape table only: a zero-stride view, no memory, no RNG
        return np.broadcast_to(np.zeros((), np.float32), shape)
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    # generate in float32 blocks to bound memory for the 131072 x 1024 Dense kernels
    out = rng.random(size=shape, dtype=np.float32)
    out *= np.float32(2.0 * limit)
    out -= np.float32(limit)
    return out


def _bias(rng, n, scale=0.05):
    return ((rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(scale)).astype(np.float32)


def _bn(rng, w: Weights, prefix: str, c: int, randomise: bool):
    if randomise:
        w[prefix + "/gamma"] = (0.5 + rng.random(c, dtype=np.float32)).astype(np.float32)            # [0.5, 1.5)
        w[prefix + "/beta"] = ((rng.random(c, dtype=np.float32) - 0.5) * np.float32(0.4)).astype(np.float32)
        w[prefix + "/moving_mean"] = ((rng.random(c, dtype=np.float32) - 0.5) * np.float32(0.4)).astype(np.float32)
        w[prefix + "/moving_variance"] = (0.5 + 1.5 * rng.random(c, dtype=np.float32)).astype(np.float32)  # [0.5, 2)
    else:
        w[prefix + "/gamma"] = np.ones(c, np.float32)
        w[prefix + "/beta"] = np.zeros(c, np.float32)
        w[prefix + "/moving_mean"] = np.zeros(c, np.float32)
        w[prefix + "/moving_variance"] = np.ones(c, np.float32)


def bn_scale_shift(w: Weights, prefix: str) -> Tuple[np.ndarray, np.ndarray]:
    """Fold inference BatchNorm into y = x*scale + shift (float64 math, float32 result)."""
    g = w[prefix + "/gamma"].astype(np.float64)
    b = w[prefix + "/beta"].astype(np.float64)
    m = w[prefix + "/moving_mean"].astype(np.float64)
    v = w[prefix + "/moving_variance"].astype(np.float64)
    s = g / np.sqrt(v + BN_EPS)
    return s.astype(np.float32), (b - m * s).astype(np.float32)


# --------------------------------------------------------------------------------------------
# autoencoder (`train_autoencoder.py:9-40`)
# --------------------------------------------------------------------------------------------
AE_LAYERS = (  # name, Cin, Cout
    ("conv1", 3, 32), ("conv2", 32, 64), ("conv3", 64, 64), ("conv_x2", 64, 64),
    ("conv5", 128, 32), ("conv_x1", 32, 32), ("conv_out", 64, 3),
)


def synthetic_autoencoder(seed: int = 42, channels: int = 3, keras_default: bool = False) -> Weights:
    rng = np.random.Generator(np.random.PCG64(seed))
    w: Weights = {}
    for name, cin, cout in AE_LAYERS:
        cin = channels if name == "conv1" else cin
        cout = channels if name == "conv_out" else cout
        w[name + "/kernel"] = _glorot(rng, (3, 3, cin, cout))
        w[name + "/bias"] = np.zeros(cout, np.float32) if keras_default else _bias(rng, cout)
    if not keras_default:
        w["conv_out/kernel"] *= np.float32(6.0)   # widen the sigmoid's input range (random init is ~flat grey)
    return w


# --------------------------------------------------------------------------------------------
# GAN codec components (`GAN_functions.py:210-331, 333-374, 495-557`)
# --------------------------------------------------------------------------------------------
def encoder_feature_dim(img_shape) -> int:
    h, w = img_shape[0], img_shape[1]
    if h % 16 or w % 16:
        raise ValueError(f"encoder needs H, W divisible by 16, got {img_shape}")
    return (h // 16) * (w // 16) * 512


def synthetic_encoder(img_shape, latent_dim: int, add_attention: bool, seed: int,
                      keras_default: bool = False, latent_std: float = 2.0) -> Weights:
    rng = np.random.Generator(np.random.PCG64(seed))
    w: Weights = {}
    chans = [img_shape[2], 64, 128, 256, 512]
    for i in range(4):
        name = f"conv{i + 1}"
        w[name + "/kernel"] = _glorot(rng, (4, 4, chans[i], chans[i + 1]))
        w[name + "/bias"] = np.zeros(chans[i + 1], np.float32) if keras_default else _bias(rng, chans[i + 1])
        if i > 0:
            _bn(rng, w, f"bn{i + 1}", chans[i + 1], randomise=not keras_default)
    if add_attention:
        for nm, co in (("query", 32), ("key", 32), ("value", 256)):
            w[f"attn/{nm}/kernel"] = _glorot(rng, (1, 1, 256, co))
            w[f"attn/{nm}/bias"] = np.zeros(co, np.float32) if keras_default else _bias(rng, co)
        w["attn/gamma"] = np.array([0.0 if keras_default else 0.5], np.float32)
    feat = encoder_feature_dim(img_shape)
    k = _glorot(rng, (feat, latent_dim))
    if not keras_default:
        # Default init gives latent std ~ 0.006 (all symbols round to 0).  Rescale the Dense
        # kernel so the latent std is ~ latent_std: each latent is a sum of `feat` products of a
        # U(+-limit) weight and an O(0.3)-RMS feature, so std ~ limit/sqrt(3) * sqrt(feat) * 0.3.
        limit = np.sqrt(6.0 / (feat + latent_dim))
        est = limit / np.sqrt(3.0) * np.sqrt(feat) * 0.15
        k *= np.float32(latent_std / est)
    w["dense/kernel"] = k
    w["dense/bias"] = np.zeros(latent_dim, np.float32) if keras_default else _bias(rng, latent_dim, 0.1)
    return w


def synthetic_generator(latent_dim: int, img_shape, seed: int, keras_default: bool = False) -> Weights:
    rng = np.random.Generator(np.random.PCG64(seed))
    w: Weights = {}
    feat = encoder_feature_dim(img_shape)
    w["dense/kernel"] = _glorot(rng, (latent_dim, feat))
    w["dense/bias"] = np.zeros(feat, np.float32) if keras_default else _bias(rng, feat)
    _bn(rng, w, "bn0", 512, randomise=not keras_default)
    # Conv2DTranspose kernels: (kh, kw, Cout, Cin); Cin includes the concatenated skip
    specs = [(512, 256), (512, 128), (256, 64), (128, 32)]
    for i, (cin, cout) in enumerate(specs, start=1):
        w[f"deconv{i}/kernel"] = _glorot(rng, (4, 4, cout, cin))
        w[f"deconv{i}/bias"] = np.zeros(cout, np.float32) if keras_default else _bias(rng, cout)
        _bn(rng, w, f"bn{i}", cout, randomise=not keras_default)
    w["conv_out/kernel"] = _glorot(rng, (4, 4, 32, img_shape[2]))
    if not keras_default:
        w["conv_out/kernel"] *= np.float32(3.0)   # widen the tanh's input range
    w["conv_out/bias"] = np.zeros(img_shape[2], np.float32) if keras_default else _bias(rng, img_shape[2])
    return w


def synthetic_latent_saliency(latent_dim: int, seed: int, keras_default: bool = False) -> Weights:
    rng = np.random.Generator(np.random.PCG64(seed))
    w: Weights = {}
    dims = [latent_dim, 512, 256, 1]
    for i in range(3):
        w[f"dense{i + 1}/kernel"] = _glorot(rng, (dims[i], dims[i + 1]))
        w[f"dense{i + 1}/bias"] = np.zeros(dims[i + 1], np.float32) if keras_default else _bias(rng, dims[i + 1])
    return w


def synthetic_rd_optimizer(seed: int, keras_default: bool = False) -> Weights:
    rng = np.random.Generator(np.random.PCG64(seed))
    w: Weights = {}
    w["conv1/kernel"] = _glorot(rng, (3, 3, 1, 32))
    w["conv1/bias"] = np.zeros(32, np.float32) if keras_default else _bias(rng, 32)
    w["conv2/kernel"] = _glorot(rng, (3, 3, 32, 64))
    w["conv2/bias"] = np.zeros(64, np.float32) if keras_default else _bias(rng, 64)
    w["dense1/kernel"] = _glorot(rng, (65, 128))
    w["dense1/bias"] = np.zeros(128, np.float32) if keras_default else _bias(rng, 128)
    w["dense2/kernel"] = _glorot(rng, (128, 3))
    w["dense2/bias"] = np.zeros(3, np.float32) if keras_default else _bias(rng, 3)
    return w


def synthetic_adaptive(img_shape, base_latent_dim: int, seed: int = 42,
                       keras_default: bool = False) -> Dict[str, Weights]:
    """Weights for the seven sub-models of `build_adaptive_compression_model` (`:585-600`)."""
    return {
        "hq_encoder": synthetic_encoder(img_shape, base_latent_dim * 2, True, seed + 1, keras_default),
        "hq_generator": synthetic_generator(base_latent_dim * 2, img_shape, seed + 2, keras_default),
        "lq_encoder": synthetic_encoder(img_shape, base_latent_dim, False, seed + 3, keras_default),
        "lq_generator": synthetic_generator(base_latent_dim, img_shape, seed + 4, keras_default),
        "latent_saliency_hq": synthetic_latent_saliency(base_latent_dim * 2, seed + 5, keras_default),
        "latent_saliency_lq": synthetic_latent_saliency(base_latent_dim, seed + 6, keras_default),
        "rd_optimizer": synthetic_rd_optimizer(seed + 7, keras_default),
    }


# ---- checkpoints -----------------------------------------------------------------------------------------------------------
# The reference saves each sub-model with Keras `model.save("*.h5")` (GAN_train.py:548-581) and reloads it with
# `keras.models.load_model` (GAN_test.py:37-78).  keras_h5.py reads those files directly; the flat .npz below (Keras layouts,
# "<sub_model>/<layer>/<tensor>") is the second route: `tools/convert_keras_h5.py` writes it in the reference's own environment.

def flatten(nested: Dict[str, Weights]) -> Weights:
    return {f"{sub}/{k}": np.asarray(v) for sub, ws in nested.items() for k, v in ws.items()}


def unflatten(flat: Weights) -> Dict[str, Weights]:
    nested: Dict[str, Weights] = {}
    for key, v in flat.items():
        sub, _, name = key.partition("/")
        if not name:
            raise ValueError(f"checkpoint key '{key}' is not '<sub_model>/<layer>/<tensor>'")
        nested.setdefault(sub, {})[name] = np.ascontiguousarray(v, dtype=np.float32)
    return nested


def save_npz(path: str, nested: Dict[str, Weights]) -> None:
    np.savez(path, **flatten(nested))


def load_npz(path: str) -> Dict[str, Weights]:
    with np.load(path) as f:
        return unflatten({k: f[k] for k in f.files})


def adaptive_shapes(img_shape, base_latent_dim: int) -> Dict[str, Weights]:
    """{sub_model: {name: zero-stride array of the expected shape}} of an adaptive-model checkpoint, without allocating or
    drawing the ~1.7 GB of Keras-default weights (the builders run with their kernel initialiser replaced by a shape stub)."""
    global _SHAPES_ONLY
    _SHAPES_ONLY = True
    try:
        return synthetic_adaptive(img_shape, base_latent_dim, keras_default=True)
    finally:
        _SHAPES_ONLY = False


def check_adaptive(nested: Dict[str, Weights], img_shape, base_latent_dim: int) -> None:
    """Raise ValueError naming the first missing / extra / mis-shaped tensor of an adaptive-model checkpoint."""
    want = adaptive_shapes(img_shape, base_latent_dim)
    for sub, ws in want.items():
        if sub not in nested:
            raise ValueError(f"checkpoint has no sub-model '{sub}'")
        for k, v in ws.items():
            if k not in nested[sub]:
                raise ValueError(f"checkpoint is missing '{sub}/{k}' {v.shape}")
            if tuple(nested[sub][k].shape) != tuple(v.shape):
                raise ValueError(f"'{sub}/{k}' has shape {tuple(nested[sub][k].shape)}, expected {tuple(v.shape)}")
        extra = set(nested[sub]) - set(ws)
        if extra:
            raise ValueError(f"checkpoint has unknown tensors in '{sub}': {sorted(extra)[:4]}")
    extra = set(nested) - set(want)
    if extra:
        raise ValueError(f"checkpoint has unknown sub-models: {sorted(extra)}")
