"""Golden vectors produced by the REFERENCE ITSELF (tools/make_reference_fixtures.py, run in a TensorFlow environment).

The files cannot be generated in this repository's container (no TensorFlow / scikit-image), so every test here skips with a
message until someone commits them; from then on they pin the oracle (CPU test) and the CUDA path (GPU test) to the reference's
own outputs with the north-star criteria."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed: run tools/make_reference_fixtures.py in the reference's TensorFlow environment")
    return np.load(path)


def _inputs(cic_mods, g):
    synth, W = cic_mods
    seed_x = int(g["seed_inputs"])
    img = synth.to_signed_range(synth.synth_images_u8(2, 256, 256, seed=seed_x))
    mask = synth.synth_masks(2, 256, 256, seed=seed_x)
    ws = W.synthetic_adaptive((256, 256, 3), 512, seed=int(g["seed_weights"]))
    return ws, img, mask, g["bpp"]


@pytest.fixture(scope="module")
def pure_mods():
    import importlib
    return (importlib.import_module("contextual-image-compression_b200.synth"), importlib.import_module("contextual-image-compression_b200.weights"))


def test_oracle_matches_reference_adaptive(pure_mods):
    from oracle import graphs, parity
    g = _load("reference_adaptive.npz")
    ws, img, mask, bpp = _inputs(pure_mods, g)
    outs, ex = graphs.adaptive_forward(ws, img, mask, bpp, return_extras=True)
    assert np.abs(ex["hq_latent"] - g["hq_latent"]).max() < 2e-4            # two fp32 implementations of a K = 131072 dot product
    assert np.abs(ex["lq_latent"] - g["lq_latent"]).max() < 2e-4
    # symbols: the reference's symbol = round(deq * scale); band criterion around the oracle's pre-round value
    for br, k in (("hq", 1), ("lq", 2)):
        ref_sym = np.rint(g[f"{br}_latent_q"].astype(np.float64) * ex[f"{br}_scale"].reshape(-1, 1))
        bad, outside = parity.symbol_parity(ex[f"{br}_sym"], ref_sym, ex[f"{br}_pre"])
        assert outside == 0, (br, bad, outside)
    np.testing.assert_allclose(outs[4], g["dt"], atol=2e-6)
    np.testing.assert_allclose(outs[3], g["rd_params"], atol=2e-5)
    assert np.abs(outs[0] - g["blended"]).max() / 2 < 1e-3


def test_oracle_matches_reference_metrics():
    from oracle import metrics
    g = _load("reference_metrics.npz")
    for a, b, row in zip(g["a8"], g["b8"], g["rows"]):
        assert metrics.ae_calculate_mse(a, b) == pytest.approx(row[0], abs=1e-9)
        assert metrics.ae_calculate_psnr(a, b) == pytest.approx(row[1], abs=1e-9)
        assert metrics.ae_calculate_ssim(a, b) == pytest.approx(row[2], abs=1e-9)
    ga = _load("reference_adaptive.npz")
    # compute_metrics (GAN_functions.py:724-759) on the reference's own reconstruction
    import importlib
    synth = importlib.import_module("contextual-image-compression_b200.synth")
    img = synth.to_signed_range(synth.synth_images_u8(2, 256, 256, seed=int(ga["seed_inputs"])))
    for i in range(2):
        m = metrics.compute_metrics(img[i], ga["blended"][i])
        assert m["psnr"] == pytest.approx(ga["metrics"][i, 0], abs=1e-4) and m["ssim"] == pytest.approx(ga["metrics"][i, 1], abs=1e-5)


def test_oracle_matches_reference_autoencoder(pure_mods):
    from oracle import graphs
    g = _load("reference_autoencoder.npz")
    synth, W = pure_mods
    x = synth.to_unit_range(synth.synth_images_u8(2, 64, 96, seed=int(g["seed_inputs"])))
    y = graphs.autoencoder_forward(W.synthetic_autoencoder(seed=int(g["seed_weights"])), x)
    assert np.abs(y - g["y"]).max() < 2e-5


@pytest.mark.gpu
def test_cuda_path_matches_reference_adaptive(cic):
    from oracle import parity
    g = _load("reference_adaptive.npz")
    ws, img, mask, bpp = _inputs((cic.synth, cic.weights), g)
    import GAN_functions as gf
    models = gf.build_adaptive_compression_model((256, 256, 3), 512, target_bpp=True)
    models["adaptive_model"].set_weights_dict(ws)
    cic.set_precision("tc")
    out = models["adaptive_model"].forward_device([cic.runtime.to_device_f32(a) for a in (img, mask, bpp)], extras=True)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    same = np.ones(2, bool)
    for br in ("hq", "lq"):
        ref_sym = np.rint(g[f"{br}_latent_q"].astype(np.float64) * out[f"{br}_scale"].reshape(-1, 1))
        pre = g[f"{br}_latent"].astype(np.float64) * out[f"{br}_scale"].reshape(-1, 1)
        bad, outside = parity.symbol_parity(out[f"{br}_symbols"], ref_sym, pre)
        assert outside == 0, (br, bad, outside)
        same &= (out[f"{br}_symbols"] == ref_sym).all(axis=1)
    assert same.any()
    assert np.abs(out["blended"][same] - g["blended"][same]).max() / 2 < 1e-2
    np.testing.assert_allclose(out["dt"], g["dt"], atol=3e-6)
    for i in np.flatnonzero(same):
        m = gf.compute_metrics(img[i], out["blended"][i])
        assert abs(m["psnr"] - g["metrics"][i, 0]) < 0.05


def _saliency_inputs(synth, g):
    return synth.to_signed_range(synth.synth_images_u8(3, 256, 256, seed=int(g["seed_inputs"])))


def test_oracle_matches_reference_saliency(pure_mods):
    """oracle/saliency.py (restated from the opencv-contrib source) against the reference's own compute_saliency_map outputs."""
    from oracle import saliency as osal
    g = _load("reference_saliency.npz")
    img = _saliency_inputs(pure_mods[0], g)
    for method, tol in (("spectral_residual", 1e-3), ("fine_grained", 1.5 / 255), ("combined", 2e-3)):
        for i in range(3):
            np.testing.assert_allclose(osal.compute_saliency_map(img[i], method), g[method][i], atol=tol)


@pytest.mark.gpu
def test_cuda_path_matches_reference_saliency(cic):
    g = _load("reference_saliency.npz")
    img = _saliency_inputs(cic.synth, g)
    for method, tol in (("spectral_residual", 1e-3), ("fine_grained", 1.5 / 255), ("combined", 2e-3)):
        np.testing.assert_allclose(cic.ops.saliency_map(img, method).cpu().numpy(), g[method], atol=tol)
    np.testing.assert_allclose(cic.ops.saliency_mask_from_image(img).cpu().numpy(), g["masks"], atol=2e-3)
