"""The C-ABI library loads on a CPU-only box, exports every symbol include/cic.h declares, and fails
loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cic.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cic_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(cic):
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(cic._lib.lib, n), f"libcic.so does not export {n}"
    assert sorted(cic._lib.PROTOTYPES) == names, "ctypes prototypes and include/cic.h disagree"


def test_version_and_error_string(cic):
    assert cic._lib.lib.cic_version() == 100
    assert isinstance(cic._lib.last_error(), str)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "contextual-image-compression_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt:
                    offenders.append(f)
    for f in ("GAN_functions.py", "GAN_test.py", "train_autoencoder.py", "test_autoencoder.py", "cic_b200.py"):
        txt = open(os.path.join(ROOT, f)).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
            offenders.append(f)
    assert not offenders, f"product code must not use the oracle: {offenders}"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(cic):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cic.ops.rate_scalars([1.0])
    m = cic.autoencoder.build_autoencoder((32, 32, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.predict(np.zeros((1, 32, 32, 3), np.float32))
    opts = cic._lib.cic_plan_opts()
    h = cic._lib.lib.cic_plan_create(cic._lib.PLAN_AUTOENCODER, None, 0, ctypes.byref(opts))
    assert not h and "no CUDA device" in cic._lib.last_error()


def test_argument_validation_without_gpu(cic):
    lib = cic._lib.lib
    assert lib.cic_quantize_latent(None, None, None, None, None, None, None, 1, 8, None) == cic._lib.ERR_INVALID
    assert "null" in cic._lib.last_error()
    assert lib.cic_quantize_latent(None, None, None, None, None, None, None, 0, 8, None) == cic._lib.OK   # empty batch
    assert lib.cic_hq_ratio_sweep(1, 1, 99, 1, 1, 16, None) == cic._lib.ERR_INVALID
    assert lib.cic_metrics_psnr_ssim_f32(1, 1, 1, 1, 4, 4, 3, 0.0, 1.0, 1.0, None) == cic._lib.ERR_INVALID
    assert "7x7" in cic._lib.last_error()
    assert lib.cic_plan_workspace_bytes(None, 4, 8, 8) == 0
