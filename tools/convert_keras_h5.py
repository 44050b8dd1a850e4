#!/usr/bin/env python
"""Convert the reference's Keras checkpoints to the flat .npz this repository loads.

Run it ONCE in the reference's own environment (TensorFlow / Keras + h5py, with the reference's GAN_functions.py importable):

    python tools/convert_keras_h5.py --model-dir models --out models/adaptive_weights.npz

It loads the seven component models the reference saves at the end of training (GAN_train.py:569-581:
`hq_encoder_final.h5`, `hq_generator_final.h5`, `lq_encoder_final.h5`, `lq_generator_final.h5`, `latent_saliency_hq_final.h5`,
`latent_saliency_lq_final.h5`, `rd_optimizer_final.h5`) with `keras.models.load_model` exactly as GAN_test.py:37-78 does, walks
`model.layers` in creation order and names the tensors the way contextual-image-compression_b200/weights.py does; layouts are
Keras' own (Conv2D (kh,kw,Cin,Cout), Conv2DTranspose (kh,kw,Cout,Cin), Dense (in,out)), so nothing is transposed.

Since the package reads .h5 files itself (contextual-image-compression_b200/keras_h5.py on hdf5_lite.py, `GAN_test.load_models(model_dir)`), this
script is the second route: it goes through Keras' own loader, so it also works for checkpoints in formats hdf5_lite does not
cover, and the two routes can be compared on a real checkpoint (`--check` reads the same files with hdf5_lite and asserts equality).

THIS SCRIPT CANNOT BE EXECUTED IN THIS REPOSITORY'S CONTAINER (no TensorFlow, no h5py): the Keras-loading part is unverified here.
The layer-to-name mapping (`map_layers`, shared with keras_h5.py) is pure Python on duck-typed layers and is covered by
tests/test_host_logic.py; it follows the creation order of GAN_functions.py:236-331 (generator, encoder), :210-234 (latent
saliency), :495-557 (RD optimizer).
"""
from __future__ import annotations

import argparse
import os

import numpy as np


def _keras_h5():
    """contextual-image-compression_b200/keras_h5.py (+ hdf5_lite.py) imported under a shell package: both are pure Python + numpy, whereas the
    package's own __init__ loads the CUDA library"""
    import importlib
    import sys
    import types
    pkg = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "contextual-image-compression_b200")
    shell = types.ModuleType("cic_h5pkg")
    shell.__path__ = [pkg]
    sys.modules.setdefault("cic_h5pkg", shell)
    return importlib.import_module("cic_h5pkg.keras_h5")


_K = _keras_h5()
map_layers, SUB_MODELS = _K.map_layers, _K.SUB_MODELS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-dir", default="models")
    ap.add_argument("--suffix", default="_final.h5", help="file name suffix of the component checkpoints (e.g. _epoch_50.h5)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--check", action="store_true", help="also read the files with the package's own HDF5 reader and compare")
    args = ap.parse_args()
    from tensorflow import keras                                   # noqa: PLC0415  (only available in the reference's environment)
    from GAN_functions import SelfAttention, AdaptiveQuantizationLayer  # the reference's module, GAN_test.py:14-20
    custom = {"SelfAttention": SelfAttention, "AdaptiveQuantizationLayer": AdaptiveQuantizationLayer}
    flat = {}
    for sub, kind in SUB_MODELS:
        model = keras.models.load_model(os.path.join(args.model_dir, sub + args.suffix), custom_objects=custom, compile=False)
        for k, v in map_layers(list(model.layers), kind).items():
            flat[f"{sub}/{k}"] = v
    if args.check:
        own = _K.load_adaptive_dir(args.model_dir, args.suffix)
        for sub, ws in own.items():
            for k, v in ws.items():
                if not np.array_equal(v, flat[f"{sub}/{k}"]):
                    raise SystemExit(f"hdf5_lite and Keras disagree on {sub}/{k}")
        print(f"hdf5_lite read the same {sum(len(w) for w in own.values())} tensors bit for bit")
    out = args.out or os.path.join(args.model_dir, "adaptive_weights.npz")
    np.savez(out, **flat)
    print(f"wrote {len(flat)} tensors, {sum(v.size for v in flat.values()):,} parameters -> {out}")


if __name__ == "__main__":
    main()
