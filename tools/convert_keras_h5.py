#!/usr/bin/env python
"""Convert the reference's Keras checkpoints to the flat .npz this repository loads.

Run it ONCE in the reference's own environment (TensorFlow / Keras + h5py, with the reference's GAN_functions.py importable):

    python tools/convert_keras_h5.py --model-dir models --out models/adaptive_weights.npz

It loads the seven component models the reference saves at the end of training (GAN_train.py:569-581:
`hq_encoder_final.h5`, `hq_generator_final.h5`, `lq_encoder_final.h5`, `lq_generator_final.h5`, `latent_saliency_hq_final.h5`,
`latent_saliency_lq_final.h5`, `rd_optimizer_final.h5`) with `keras.models.load_model` exactly as GAN_test.py:37-78 does, walks
`model.layers` in creation order and names the tensors the way contextual-image-compression_b200/weights.py does; layouts are
Keras' own (Conv2D (kh,kw,Cin,Cout), Conv2DTranspose (kh,kw,Cout,Cin), Dense (in,out)), so nothing is transposed.

THIS SCRIPT CANNOT BE EXECUTED IN THIS REPOSITORY'S CONTAINER (no TensorFlow, no h5py): the Keras-loading part is unverified here.
The layer-to-name mapping (`map_layers`) is pure Python on duck-typed layers and is covered by tests/test_host_logic.py; it follows
the creation order of GAN_functions.py:236-331 (generator, encoder), :210-234 (latent saliency), :495-557 (RD optimizer).
"""
from __future__ import annotations

import argparse
import os

import numpy as np


def _cls(layer) -> str:
    return type(layer).__name__


def map_layers(layers, kind: str) -> dict:
    """{name: array} for one sub-model from its layers in creation order (objects with a class name and get_weights())."""
    out = {}
    convs = [l for l in layers if _cls(l) == "Conv2D"]
    deconvs = [l for l in layers if _cls(l) == "Conv2DTranspose"]
    denses = [l for l in layers if _cls(l) == "Dense"]
    bns = [l for l in layers if _cls(l) == "BatchNormalization"]
    attn = [l for l in layers if _cls(l) == "SelfAttention"]

    def put(prefix, layer, names=("kernel", "bias")):
        ws = layer.get_weights()
        if len(ws) != len(names):
            raise ValueError(f"{kind}/{prefix}: expected {len(names)} tensors, found {len(ws)}")
        for n, w in zip(names, ws):
            out[f"{prefix}/{n}"] = np.asarray(w, np.float32)

    bn_names = ("gamma", "beta", "moving_mean", "moving_variance")
    if kind == "encoder":                      # GAN_functions.py:300-326: conv1, (conv + BN) x 3, [attention before conv4], Dense
        if len(convs) != 4 or len(bns) != 3 or len(denses) != 1 or len(attn) > 1:
            raise ValueError(f"encoder: unexpected layer counts conv={len(convs)} bn={len(bns)} dense={len(denses)} attn={len(attn)}")
        for i, l in enumerate(convs, start=1):
            put(f"conv{i}", l)
        for i, l in enumerate(bns, start=2):
            put(f"bn{i}", l, bn_names)
        put("dense", denses[0])
        if attn:                               # :339-342: weights in creation order gamma, then query / key / value kernel + bias
            a = attn[0]
            for nm, sub in (("query", a.query_conv), ("key", a.key_conv), ("value", a.value_conv)):
                put(f"attn/{nm}", sub)
            out["attn/gamma"] = np.asarray(a.gamma.numpy() if hasattr(a.gamma, "numpy") else a.gamma, np.float32).reshape(1)
    elif kind == "generator":                  # :247-273: Dense, BN, (ConvT + BN) x 4, Conv2D
        if len(denses) != 1 or len(bns) != 5 or len(deconvs) != 4 or len(convs) != 1:
            raise ValueError(f"generator: unexpected layer counts dense={len(denses)} bn={len(bns)} deconv={len(deconvs)} conv={len(convs)}")
        put("dense", denses[0])
        for i, l in enumerate(bns):
            put(f"bn{i}", l, bn_names)
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
    else:
        raise ValueError(f"unknown sub-model kind '{kind}'")
    return out


SUB_MODELS = (("hq_encoder", "encoder"), ("hq_generator", "generator"), ("lq_encoder", "encoder"), ("lq_generator", "generator"),
              ("latent_saliency_hq", "latent_saliency"), ("latent_saliency_lq", "latent_saliency"), ("rd_optimizer", "rd_optimizer"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-dir", default="models")
    ap.add_argument("--suffix", default="_final.h5", help="file name suffix of the component checkpoints (e.g. _epoch_50.h5)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from tensorflow import keras                                   # noqa: PLC0415  (only available in the reference's environment)
    from GAN_functions import SelfAttention, AdaptiveQuantizationLayer  # the reference's module, GAN_test.py:14-20
    custom = {"SelfAttention": SelfAttention, "AdaptiveQuantizationLayer": AdaptiveQuantizationLayer}
    flat = {}
    for sub, kind in SUB_MODELS:
        model = keras.models.load_model(os.path.join(args.model_dir, sub + args.suffix), custom_objects=custom, compile=False)
        for k, v in map_layers(list(model.layers), kind).items():
            flat[f"{sub}/{k}"] = v
    out = args.out or os.path.join(args.model_dir, "adaptive_weights.npz")
    np.savez(out, **flat)
    print(f"wrote {len(flat)} tensors, {sum(v.size for v in flat.values()):,} parameters -> {out}")


if __name__ == "__main__":
    main()
