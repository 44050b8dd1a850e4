"""The oracle against the closed-form known answers and the published plot band."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import graphs, metrics

HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))


def test_rate_scalars():
    for row in KA["rate_scalars"]:
        t, thr, qs = graphs.rate_scalars([row["bpp"]])
        assert abs(float(t) - row["t"]) < 2e-6
        assert abs(float(thr) - row["thr"]) < 2e-6
        assert abs(float(qs) - row["qs"]) < 2e-6


def test_dynamic_threshold():
    m = np.array(KA["dynamic_threshold"]["mask"], np.float32).reshape(1, -1, 1, 1)
    for bpp in ("0.1", "1.0", "2.0"):
        dt = graphs.dynamic_threshold(m, [float(bpp)]).ravel()
        np.testing.assert_allclose(dt, KA["dynamic_threshold"][bpp], atol=2e-6)


def test_quantizer_known_answers_and_half_even():
    for row in KA["quantizer"]:
        _, _, qs = graphs.rate_scalars([row["bpp"]])
        deq, sym, pre, scale = graphs.adaptive_quantize(np.array([[row["latent"]]], np.float32),
                                                        np.array([[row["sal"]]], np.float32), qs)
        assert abs(float(scale) - row["scale"]) < 1e-5 * max(1, row["scale"])
        assert abs(float(pre) - row["pre"]) < 2e-5 * max(1, abs(row["pre"]))
        assert int(sym) == row["symbol"]
        assert abs(float(deq) - row["deq"]) < 2e-6 * max(1, abs(row["deq"]))


def test_bpp_accounting():
    for row in KA["bpp_accounting"]:
        acc = metrics.bpp_accounting(np.full((256, 256, 1), row["hq_ratio"], np.float64))
        assert abs(acc["total_bits"] - row["total_bits"]) < 1e-6
        assert abs(acc["actual_bpp"] - row["actual_bpp"]) < 1e-9
        assert abs(acc["compression_ratio"] - row["compression_ratio"]) < 1e-9
        assert abs(acc["actual_bpp"] - 0.25 * (1 + row["hq_ratio"])) < 1e-12


def test_psnr_identities():
    for row in KA["psnr_identities"]:
        a = np.full((16, 16, 3), 0.5, np.float32)
        b = a + np.float32(row["uniform_error"])
        assert abs(metrics.sk_psnr(a, b, row["data_range"]) - row["psnr_db"]) < 2e-3
    assert metrics.sk_psnr(a, a, 1.0) == float("inf")


def test_hq_ratio_in_published_band_and_monotone():
    """hq_ratio_by_bpp.png: monotone increasing in target bpp; 0.017-0.061 at 0.1 bpp, 0.068-0.199 at 2.0."""
    import importlib
    synth = importlib.import_module("contextual-image-compression_b200.synth")
    masks = synth.synth_masks(8, 256, 256)
    levels = synth.rate_control_bpps()
    ratios = np.stack([graphs.dynamic_threshold(masks, np.full(8, b, np.float32)).reshape(8, -1).mean(1) for b in levels], 1)
    assert np.all(np.diff(ratios, axis=1) > 0)
    assert ratios[:, 0].min() > 0.01 and ratios[:, 0].max() < 0.07
    assert ratios[:, -1].min() > 0.06 and ratios[:, -1].max() < 0.25
