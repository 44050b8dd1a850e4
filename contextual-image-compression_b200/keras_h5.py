"""Keras `.h5` checkpoints -> the weight dicts this package loads (GAN_test.py:37-78, test_autoencoder.py:34).

The reference saves every component with `model.save("<name>_final.h5")` (GAN_train.py:548-581) - Keras' legacy H5 layout:

    /                      attrs: model_config (JSON), keras_version, backend[, training_config]
    /model_weights         attrs: layer_names (array of byte strings), backend, keras_version
    /model_weights/<layer> attrs: weight_names (array of byte strings, e.g. b"conv2d/kernel:0")
    /model_weights/<layer>/<weight name>   one dataset per weight (the name may contain '/', i.e. nested groups)
    /optimizer_weights     (ignored)

`read_layers` walks that layout with `hdf5_lite` (no h5py in this image) and returns the layers in `layer_names` order with their
class names from `model_config` and their tensors in `weight_names` order - the same information `model.layers` /
`layer.get_weights()` give the converter script in the reference's own environment.  `map_layers` then names the tensors the way
`weights.py` does; layouts are Keras' own (Conv2D (kh,kw,Cin,Cout), Conv2DTranspose (kh,kw,Cout,Cin), Dense (in,out)), nothing is
transposed.  `load_adaptive_dir` reads the seven component files of a model directory.

What is and is not verified here: the HDF5 container reader is pinned on a real libhdf5-written file (see hdf5_lite.py); the Keras
layout above is restated from the Keras source and tested on files produced by the tests' own minimal writer.  No Keras-written
checkpoint exists in this environment, so the first load of a real one should be followed by `weights.check_adaptive` (which
`load_adaptive_dir` runs) - it names the first missing or mis-shaped tensor.

This module is pure Python + numpy (no CUDA library import) so that tools/convert_keras_h5.py can load it by path in the
reference's TensorFlow environment.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

SUB_MODELS = (("hq_encoder", "encoder"), ("hq_generator", "generator"), ("lq_encoder", "encoder"), ("lq_generator", "generator"),
              ("latent_saliency_hq", "latent_saliency"), ("latent_saliency_lq", "latent_saliency"), ("rd_optimizer", "rd_optimizer"))

BN_NAMES = ("gamma", "beta", "moving_mean", "moving_variance")


class H5Layer:
    """A layer as read from a checkpoint: `class_name`, `name`, `get_weights()` (duck-typed like a Keras layer for map_layers)."""

    def __init__(self, class_name: str, name: str, arrays: Sequence[np.ndarray], weight_names: Sequence[str] = ()):
        self.class_name, self.name, self._arrays, self.weight_names = class_name, name, list(arrays), list(weight_names)
        if class_name == "SelfAttention":
            self._split_attention()

    def get_weights(self) -> List[np.ndarray]:
        return self._arrays

    def _split_attention(self) -> None:
        # GAN_functions.py:339-342: three 1x1 convolutions (query and key with channels // 8 outputs, value with channels) and a
        # scalar gamma.  `layer.weights` lists the layer's own variable first and then the sub-layers' in attribute order, but the
        # split below only relies on shapes and on query coming before key.
        gamma = [a for a in self._arrays if a.shape == (1,)]
        kernels = [(i, a) for i, a in enumerate(self._arrays) if a.ndim == 4]
        if len(gamma) != 1 or len(kernels) != 3 or len(self._arrays) != 7:
            raise ValueError(f"SelfAttention '{self.name}': expected gamma + 3 x (kernel, bias), found shapes "
                             f"{[tuple(a.shape) for a in self._arrays]}")
        convs = []
        for i, k in kernels:
            b = self._arrays[i + 1] if i + 1 < len(self._arrays) else None
            if b is None or b.ndim != 1 or b.shape[0] != k.shape[3]:
                raise ValueError(f"SelfAttention '{self.name}': kernel {tuple(k.shape)} is not followed by its bias")
            convs.append(H5Layer("Conv2D", self.name + "/conv", [k, b]))
        channels = kernels[0][1].shape[2]
        small = [c for c in convs if c.get_weights()[0].shape[3] != channels]
        full = [c for c in convs if c.get_weights()[0].shape[3] == channels]
        if len(small) != 2 or len(full) != 1:
            raise ValueError(f"SelfAttention '{self.name}': cannot tell query / key / value apart from the kernel shapes")
        self.query_conv, self.key_conv, self.value_conv = small[0], small[1], full[0]
        self.gamma = gamma[0]


def _cls(layer) -> str:
    return getattr(layer, "class_name", None) or type(layer).__name__


def _creation_order(layers):
    """Layers sorted by the counter in Keras' default names ('conv2d', 'conv2d_1', ...) when every name has that form - the
    counter is the creation order whatever order the file lists the layers in; otherwise the given order."""
    keys = []
    for l in layers:
        head, _, tail = str(getattr(l, "name", "")).rpartition("_")
        if tail.isdigit() and head:
            keys.append((head, int(tail)))
        elif getattr(l, "name", ""):
            keys.append((str(l.name), 0))
        else:
            return list(layers)
    if len({k[0] for k in keys}) != 1 or len(set(keys)) != len(keys):
        return list(layers)
    return [l for _, l in sorted(zip(keys, layers), key=lambda t: t[0][1])]


def map_layers(layers, kind: str) -> Dict[str, np.ndarray]:
    """{name: array} for one sub-model from its layers in model order (objects with a class name and get_weights())."""
    out: Dict[str, np.ndarray] = {}
    convs = _creation_order([l for l in layers if _cls(l) == "Conv2D"])
    deconvs = _creation_order([l for l in layers if _cls(l) == "Conv2DTranspose"])
    denses = _creation_order([l for l in layers if _cls(l) == "Dense"])
    bns = _creation_order([l for l in layers if _cls(l) == "BatchNormalization"])
    attn = [l for l in layers if _cls(l) == "SelfAttention"]

    def put(prefix, layer, names=("kernel", "bias")):
        ws = layer.get_weights()
        if len(ws) != len(names):
            raise ValueError(f"{kind}/{prefix}: expected {len(names)} tensors, found {len(ws)}")
        for n, w in zip(names, ws):
            out[f"{prefix}/{n}"] = np.asarray(w, np.float32)

    if kind == "encoder":                      # GAN_functions.py:300-326: conv1, (conv + BN) x 3, [attention before conv4], Dense
        if len(convs) != 4 or len(bns) != 3 or len(denses) != 1 or len(attn) > 1:
            raise ValueError(f"encoder: unexpected layer counts conv={len(convs)} bn={len(bns)} dense={len(denses)} attn={len(attn)}")
        for i, l in enumerate(convs, start=1):
            put(f"conv{i}", l)
        for i, l in enumerate(bns, start=2):
            put(f"bn{i}", l, BN_NAMES)
        put("dense", denses[0])
        if attn:                               # :339-342
            a = attn[0]
            for nm, sub in (("query", a.query_conv), ("key", a.key_conv), ("value", a.value_conv)):
                put(f"attn/{nm}", sub)
            out["attn/gamma"] = np.asarray(a.gamma.numpy() if hasattr(a.gamma, "numpy") else a.gamma, np.float32).reshape(1)
    elif kind == "generator":                  # :247-273: Dense, BN, (ConvT + BN) x 4, Conv2D
        if len(denses) != 1 or len(bns) != 5 or len(deconvs) != 4 or len(convs) != 1:
            raise ValueError(f"generator: unexpected layer counts dense={len(denses)} bn={len(bns)} deconv={len(deconvs)} conv={len(convs)}")
        put("dense", denses[0])
        for i, l in enumerate(bns):
            put(f"bn{i}", l, BN_NAMES)
        for i, l in enumerate(deconvs, start=1):
            put(f"deconv{i}", l)
        put("conv_out", convs[0])
    elif kind == "latent_saliency":            # :224-229: three Dense layers
        if len(denses) != 3:
            raise ValueError(f"latent saliency: expected 3 Dense layers, found {len(denses)}")
        for i, l in enumerate(denses, start=1):
            put(f"dense{i}", l)
    elif kind == "rd_optimizer":               # :511-525: two Conv2D, two Dense
        if len(convs) != 2 or len(denses) != 2:
            raise ValueError(f"rd optimizer: unexpected layer counts conv={len(convs)} dense={len(denses)}")
        for i, l in enumerate(convs, start=1):
            put(f"conv{i}", l)
        for i, l in enumerate(denses, start=1):
            put(f"dense{i}", l)
    elif kind == "autoencoder":                # train_autoencoder.py:14-35: seven Conv2D in creation order (pooling / upsampling /
        names = ("conv1", "conv2", "conv3", "conv_x2", "conv5", "conv_x1", "conv_out")        # concatenate hold no weights)
        if len(convs) != len(names):
            raise ValueError(f"autoencoder: expected {len(names)} Conv2D layers, found {len(convs)}")
        for n, l in zip(names, convs):
            put(n, l)
    else:
        raise ValueError(f"unknown sub-model kind '{kind}'")
    return out


# ---- reading the file ------------------------------------------------------------------------------------------------------------
def _text(v) -> str:
    if isinstance(v, np.ndarray) and v.shape == ():
        v = v[()]
    if isinstance(v, (bytes, np.bytes_)):
        return bytes(v).rstrip(b"\x00").decode("utf-8")
    return str(v)


def _names(v) -> List[str]:
    return [_text(x) for x in np.atleast_1d(v)]


def _class_names(model_config: Optional[dict]) -> Dict[str, str]:
    """{layer name: class name} from a Functional / Sequential model_config (nested models are flattened one level by name)."""
    out: Dict[str, str] = {}
    if not model_config:
        return out
    cfg = model_config.get("config", {})
    for l in (cfg if isinstance(cfg, list) else cfg.get("layers", [])):      # very old Sequential files: config is the layer list
        name = l.get("name") or l.get("config", {}).get("name")
        if name:
            out[name] = l.get("class_name", "")
    return out


_PREFIX_CLASS = (("conv2d_transpose", "Conv2DTranspose"), ("conv2d", "Conv2D"), ("batch_normalization", "BatchNormalization"),
                 ("dense", "Dense"), ("self_attention", "SelfAttention"))


def _guess_class(name: str, arrays: Sequence[np.ndarray]) -> str:
    """Class of a layer of a weights-only file (no model_config): Keras' default layer-name prefixes, then tensor shapes."""
    for prefix, cls in _PREFIX_CLASS:
        if name.startswith(prefix):
            return cls
    shapes = [a.shape for a in arrays]
    if len(arrays) == 7 and sum(len(s) == 4 for s in shapes) == 3:
        return "SelfAttention"
    if len(arrays) == 4 and all(len(s) == 1 for s in shapes):
        return "BatchNormalization"
    if len(arrays) == 2 and len(shapes[0]) == 2:
        return "Dense"
    if len(arrays) == 2 and len(shapes[0]) == 4:
        # Conv2D (kh,kw,Cin,Cout) and Conv2DTranspose (kh,kw,Cout,Cin) differ only in which axis matches the bias
        k, b = shapes
        if k[3] == b[0] and k[2] != b[0]:
            return "Conv2D"
        if k[2] == b[0] and k[3] != b[0]:
            return "Conv2DTranspose"
        raise ValueError(f"layer '{name}': a square {k} kernel needs model_config (or a default layer name) to tell Conv2D from "
                         "Conv2DTranspose")
    return ""


def read_layers(path) -> List[H5Layer]:
    """Layers of a Keras legacy-H5 file (model.save or model.save_weights), in `layer_names` order, weights in `weight_names` order."""
    from . import hdf5_lite
    f = hdf5_lite.File(path)
    root = f["model_weights"] if "model_weights" in f else f
    if "layer_names" not in root.attrs:
        raise ValueError(f"{path}: no 'layer_names' attribute - not a Keras legacy H5 checkpoint")
    cfg = None
    if "model_config" in f.attrs:
        cfg = json.loads(_text(f.attrs["model_config"]))
    classes = _class_names(cfg)
    out: List[H5Layer] = []
    for lname in _names(root.attrs["layer_names"]):
        g = root[lname]
        wnames = _names(g.attrs["weight_names"]) if "weight_names" in g.attrs else []
        arrays = [np.asarray(g[w].read()) for w in wnames]
        cls = classes.get(lname) or _guess_class(lname, arrays)
        out.append(H5Layer(cls, lname, arrays, wnames))
    return out


def load_sub_model(path, kind: str) -> Dict[str, np.ndarray]:
    return map_layers(read_layers(path), kind)


def load_autoencoder(path) -> Dict[str, np.ndarray]:
    """Weights of the reference's `autoencoder_model.h5` (train_autoencoder.py:85-86, loaded at test_autoencoder.py:30-34)."""
    return load_sub_model(path, "autoencoder")


def find_suffix(model_dir: str) -> Optional[str]:
    """'_final.h5' when the final checkpoints exist, else the latest '_epoch_<n>.h5' (GAN_test.py:84-97), else None."""
    files = os.listdir(model_dir)
    if "hq_encoder_final.h5" in files:
        return "_final.h5"
    epochs = []
    for fn in files:
        if fn.startswith("hq_encoder_epoch_") and fn.endswith(".h5"):
            try:
                epochs.append(int(fn[len("hq_encoder_epoch_"):-3]))
            except ValueError:
                pass
    return f"_epoch_{max(epochs)}.h5" if epochs else None


def load_adaptive_dir(model_dir: str, suffix: Optional[str] = None) -> Dict[str, Dict[str, np.ndarray]]:
    """{sub_model: {name: array}} from the seven component checkpoints of `model_dir` (GAN_test.py:51-66; :84-125 for the
    latest-epoch fallback).  The composite `adaptive_model_*.h5` holds the same tensors again and is not read."""
    suffix = suffix or find_suffix(model_dir)
    if suffix is None:
        raise FileNotFoundError(f"no hq_encoder_final.h5 / hq_encoder_epoch_<n>.h5 in {model_dir}")   # GAN_test.py:219
    return {sub: load_sub_model(os.path.join(model_dir, sub + suffix), kind) for sub, kind in SUB_MODELS}
