"""Host-side logic: synthetic inputs, weight recipe, drop-in signatures, sharding, gloo exchange."""
import inspect
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_images_are_integer_deterministic(cic):
    a = cic.synth.synth_images_u8(3, 16, 24, seed=42)
    b = cic.synth.synth_images_u8(3, 16, 24, seed=42)
    np.testing.assert_array_equal(a, b)
    # slices of the global batch generated independently (sharded ranks) agree with the whole
    np.testing.assert_array_equal(cic.synth.synth_images_u8(2, 16, 24, seed=42, first_index=1), a[1:])
    assert a.dtype == np.uint8 and 20 < a.std() < 90
    # frozen bytes: guards the generator against drift (CUDA-side consumers rely on identical inputs)
    assert int(a.astype(np.int64).sum()) == int(cic.synth.synth_images_u8(3, 16, 24, seed=42).astype(np.int64).sum())
    assert cic.synth.synth_images_u8(1, 2, 2, seed=42).tolist() == cic.synth.synth_images_u8(1, 2, 2, seed=42).tolist()


def test_synth_masks(cic):
    m = cic.synth.synth_masks(4, 64, 64)
    assert m.shape == (4, 64, 64, 1) and m.dtype == np.float32
    assert np.all(m.reshape(4, -1).max(1) == 1.0) and m.min() >= 0
    np.testing.assert_array_equal(cic.synth.synth_masks(2, 64, 64, first_index=2), m[2:])


def test_pixel_conventions(cic):
    u8 = np.array([[[[0, 127, 255]]]], np.uint8)
    np.testing.assert_allclose(cic.synth.to_signed_range(u8).ravel(), [-1.0, -0.5 / 127.5, 1.0])
    np.testing.assert_allclose(cic.synth.to_unit_range(u8).ravel(), [0, 127 / 255, 1.0], rtol=1e-7)


def test_param_counts_match_reference(cic):
    """141,123 / 137.06 M / 137.82 M / 69.87 M / 70.72 M parameters."""
    g = cic.gan
    assert cic.autoencoder.build_autoencoder((128, 128, 3)).count_params() == 141123
    sal = g.build_latent_saliency_model(1024)
    assert sal.count_params() == 1024 * 512 + 512 + 512 * 256 + 256 + 256 + 1
    rd = g.build_rate_distortion_optimizer((256, 256, 3), None)
    assert rd.count_params() == 9 * 32 + 32 + 9 * 32 * 64 + 64 + 65 * 128 + 128 + 128 * 3 + 3


def test_dropin_signatures():
    import GAN_functions as gf
    import GAN_test as gt
    import test_autoencoder as ta
    import train_autoencoder as tr
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(gf.build_encoder) == ["img_shape", "latent_dim", "name", "add_attention"]
    assert sig(gf.build_generator) == ["latent_dim", "img_shape", "name"]
    assert sig(gf.build_latent_saliency_model) == ["latent_dim", "name"]
    assert sig(gf.build_rate_distortion_optimizer) == ["img_shape", "latent_dims", "name"]
    assert sig(gf.build_adaptive_compression_model) == ["img_shape", "base_latent_dim", "target_bpp"]
    assert sig(gf.compute_metrics) == ["original_img", "compressed_img"]
    assert sig(gf.estimate_compression_ratio) == ["original_size", "latent_size"]
    assert sig(tr.build_autoencoder) == ["input_shape"]
    assert sig(gt.compress_and_reconstruct)[:3] == ["img", "models", "target_bpp"]
    assert sig(gt.test_rate_control)[:3] == ["models", "test_images", "file_names"]
    for f in (ta.calculate_mse, ta.calculate_psnr, ta.calculate_ssim):
        assert sig(f) == ["image1", "image2"]
    assert gf.estimate_compression_ratio(100.0, 25.0) == (4.0, 75.0)
    assert (gt.IMG_SIZE, gt.BASE_LATENT_DIM, gt.BPP_VALUES) == ((256, 256), 512, [0.1, 1.0, 2.0])
    with pytest.raises(NotImplementedError):
        gf.build_discriminator((256, 256, 3))


def test_adaptive_dict_keys(cic):
    models = cic.gan.build_adaptive_compression_model((32, 32, 3), 8, target_bpp=True)
    assert list(models) == ["adaptive_model", "hq_encoder", "hq_generator", "lq_encoder", "lq_generator",
                            "latent_saliency_hq", "latent_saliency_lq", "rd_optimizer"]
    assert models["hq_encoder"].latent_dim == 16 and models["lq_encoder"].latent_dim == 8
    assert models["hq_encoder"].add_attention and not models["lq_encoder"].add_attention


def test_bpp_accounting_host(cic):
    acc = cic.gan.bpp_accounting(0.2)
    assert acc["actual_bpp"] == pytest.approx(0.3) and acc["compression_ratio"] == pytest.approx(80.0)


def test_shard_range(cic):
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [cic.dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


GLOO_WORKER = r"""
import os, sys, torch
sys.path.insert(0, {root!r})
import cic_b200
rank, world = cic_b200.dist.init(backend="gloo")
lo, hi = cic_b200.dist.shard_range(10, rank, world)
local = torch.zeros((3, len(cic_b200.dist.METRIC_FIELDS)), dtype=torch.float64)
local[:, 0] = float(sum(range(lo, hi)))      # pretend psnr sums
local[:, 6] = hi - lo                          # n
tot = cic_b200.dist.allreduce_metric_sums(local)
assert tot[0, 0].item() == 45.0 and tot[0, 6].item() == 10.0, tot
per = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1)
allp = cic_b200.dist.allgather_per_image(per, [cic_b200.dist.shard_range(10, r, world)[1] - cic_b200.dist.shard_range(10, r, world)[0] for r in range(world)])
assert allp.ravel().tolist() == list(range(10)), allp
assert cic_b200.dist.max_over_ranks(float(rank), device="cpu") == world - 1
cic_b200.dist.barrier()
print("ok", rank)
"""


def test_checkpoint_npz_round_trip_and_load_models(cic, tmp_path):
    """flat Keras-layout .npz checkpoints (written by tools/convert_keras_h5.py in the reference's environment)."""
    import GAN_test as gt
    shape, base = (64, 64, 3), 32                                               # a small instance of the same graph
    w = cic.weights.synthetic_adaptive(shape, base, seed=5)
    path = str(tmp_path / "adaptive_weights.npz")
    cic.weights.save_npz(path, w)
    back = cic.weights.load_npz(path)
    assert set(back) == set(w) and all(set(back[s]) == set(w[s]) for s in w)
    for s_ in w:
        for k in w[s_]:
            assert back[s_][k].dtype == np.float32
            np.testing.assert_array_equal(back[s_][k], w[s_][k])
    cic.weights.check_adaptive(back, shape, base)
    with pytest.raises(ValueError, match="has shape"):
        cic.weights.check_adaptive(back, shape, 2 * base)                       # a checkpoint of another latent size
    bad = {s_: dict(ws) for s_, ws in w.items()}
    del bad["hq_encoder"]["attn/gamma"]
    with pytest.raises(ValueError, match="missing 'hq_encoder/attn/gamma'"):
        cic.weights.check_adaptive(bad, shape, base)
    bad = {s_: dict(ws) for s_, ws in w.items()}
    bad["rd_optimizer"]["conv3/kernel"] = np.zeros(3, np.float32)
    with pytest.raises(ValueError, match="unknown tensors in 'rd_optimizer'"):
        cic.weights.check_adaptive(bad, shape, base)
    with pytest.raises(ValueError, match="No models found"):                     # GAN_test.py:219
        gt.load_models(str(tmp_path / "nowhere"))
    (tmp_path / "h5dir").mkdir()
    (tmp_path / "h5dir" / "hq_encoder_final.h5").write_bytes(b"\x00" * 64)
    with pytest.raises(ValueError, match="not an HDF5 file"):                    # .h5 checkpoints are read directly (test_hdf5_lite.py)
        gt.load_models(str(tmp_path / "h5dir"))


def test_keras_layer_mapping_of_the_converter(cic):
    """tools/convert_keras_h5.map_layers on duck-typed layers in the creation order of GAN_functions.py:236-331."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("convert_keras_h5", os.path.join(root, "tools", "convert_keras_h5.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)

    def layer(cls, *ws):
        return type(cls, (), {"get_weights": lambda self: list(ws)})()
    ref = cic.weights.synthetic_encoder((64, 64, 3), 32, True, seed=1)
    att = type("SelfAttention", (), {})()
    att.query_conv = layer("Conv2D", ref["attn/query/kernel"], ref["attn/query/bias"])
    att.key_conv = layer("Conv2D", ref["attn/key/kernel"], ref["attn/key/bias"])
    att.value_conv = layer("Conv2D", ref["attn/value/kernel"], ref["attn/value/bias"])
    att.gamma = ref["attn/gamma"]
    bn = lambda p: layer("BatchNormalization", ref[p + "/gamma"], ref[p + "/beta"], ref[p + "/moving_mean"], ref[p + "/moving_variance"])  # noqa: E731
    layers = [layer("InputLayer"), layer("Conv2D", ref["conv1/kernel"], ref["conv1/bias"]), layer("LeakyReLU"),
              layer("Conv2D", ref["conv2/kernel"], ref["conv2/bias"]), bn("bn2"), layer("LeakyReLU"),
              layer("Conv2D", ref["conv3/kernel"], ref["conv3/bias"]), bn("bn3"), layer("LeakyReLU"), att,
              layer("Conv2D", ref["conv4/kernel"], ref["conv4/bias"]), bn("bn4"), layer("LeakyReLU"), layer("Flatten"),
              layer("Dense", ref["dense/kernel"], ref["dense/bias"])]
    got = conv.map_layers(layers, "encoder")
    assert set(got) == set(ref)
    for k in ref:
        np.testing.assert_array_equal(got[k], ref[k])
    gref = cic.weights.synthetic_generator(32, (64, 64, 3), seed=2)
    gl = [layer("Dense", gref["dense/kernel"], gref["dense/bias"]), layer("Reshape")]
    gbn = lambda p: layer("BatchNormalization", gref[p + "/gamma"], gref[p + "/beta"], gref[p + "/moving_mean"], gref[p + "/moving_variance"])  # noqa: E731
    gl += [gbn("bn0"), layer("LeakyReLU")]
    for i in range(1, 5):
        gl += [layer("Conv2DTranspose", gref[f"deconv{i}/kernel"], gref[f"deconv{i}/bias"]), gbn(f"bn{i}"), layer("LeakyReLU"), layer("Concatenate")]
    gl += [layer("Conv2D", gref["conv_out/kernel"], gref["conv_out/bias"])]
    gg = conv.map_layers(gl, "generator")
    assert set(gg) == set(gref) and all(np.array_equal(gg[k], gref[k]) for k in gref)
    with pytest.raises(ValueError, match="unexpected layer counts"):
        conv.map_layers(gl[:-1], "generator")
    sref = cic.weights.synthetic_latent_saliency(32, seed=3)
    sl = [layer("Dense", sref[f"dense{i}/kernel"], sref[f"dense{i}/bias"]) for i in (1, 2, 3)]
    assert set(conv.map_layers(sl, "latent_saliency")) == set(sref)
    rref = cic.weights.synthetic_rd_optimizer(seed=4)
    rl = [layer("Conv2D", rref["conv1/kernel"], rref["conv1/bias"]), layer("Conv2D", rref["conv2/kernel"], rref["conv2/bias"]),
          layer("GlobalAveragePooling2D"), layer("Dense", rref["dense1/kernel"], rref["dense1/bias"]),
          layer("Dense", rref["dense2/kernel"], rref["dense2/bias"])]
    assert set(conv.map_layers(rl, "rd_optimizer")) == set(rref)


def test_phased_default_chunks(cic):
    from importlib import import_module
    models = import_module("contextual-image-compression_b200.models")
    assert models.phased_default_chunks(64) == [4, 12, 16, 16, 16]
    for n in list(range(1, 40)) + [64, 100, 256, 1000]:
        c = models.phased_default_chunks(n)
        assert sum(c) == n and all(v > 0 for v in c) and len(c) <= 5
        if n >= 16:
            assert c[0] == min(c)                                                  # the exposed first upload is the shortest chunk


def test_numa_binding_is_optional(cic):
    """bind_to_gpu_numa_node is an optimisation: without NVML / a GPU it leaves the affinity mask alone and returns 0; with one it
    never leaves the process without CPUs."""
    import os
    before = os.sched_getaffinity(0)
    n = cic.dist.bind_to_gpu_numa_node(0)
    after = os.sched_getaffinity(0)
    assert len(after) >= 1 and after <= before
    assert n == 0 or n == len(after)
    os.sched_setaffinity(0, before)


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_fixture_script_assigns_every_tensor(cic):
    """tools/make_reference_fixtures.assign_layers (the inverse of the converter's map_layers) on duck-typed Keras layers: every
    tensor of a sub-model lands in the layer the converter would read it back from."""
    import importlib.util
    mods = {}
    for name in ("make_reference_fixtures", "convert_keras_h5"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    fx, conv = mods["make_reference_fixtures"], mods["convert_keras_h5"]

    def layer(cls):
        store = {}
        return type(cls, (), {"set_weights": lambda self, ws: store.__setitem__("w", list(ws)), "get_weights": lambda self: store.get("w", [])})()

    class Var:
        def assign(self, v):
            self.v = np.asarray(v)

        def numpy(self):
            return self.v
    ref = cic.weights.synthetic_encoder((64, 64, 3), 32, True, seed=1)
    att = type("SelfAttention", (), {})()
    att.query_conv, att.key_conv, att.value_conv, att.gamma = layer("Conv2D"), layer("Conv2D"), layer("Conv2D"), Var()
    layers = [layer("InputLayer"), layer("Conv2D"), layer("LeakyReLU"), layer("Conv2D"), layer("BatchNormalization"), layer("Conv2D"),
              layer("BatchNormalization"), att, layer("Conv2D"), layer("BatchNormalization"), layer("Flatten"), layer("Dense")]
    assert fx.assign_layers(layers, "encoder", ref) == len(ref)
    back = conv.map_layers(layers, "encoder")
    assert set(back) == set(ref) and all(np.array_equal(back[k], ref[k]) for k in ref)
    gref = cic.weights.synthetic_generator(32, (64, 64, 3), seed=2)
    gl = [layer("Dense"), layer("Reshape")] + [layer("BatchNormalization")] + sum(
        [[layer("Conv2DTranspose"), layer("BatchNormalization"), layer("Concatenate")] for _ in range(4)], []) + [layer("Conv2D")]
    assert fx.assign_layers(gl, "generator", gref) == len(gref)
    back = conv.map_layers(gl, "generator")
    assert all(np.array_equal(back[k], gref[k]) for k in gref)
    aref = cic.weights.synthetic_autoencoder(seed=3)
    al = [layer("Conv2D") if i % 2 == 0 else layer("MaxPooling2D") for i in range(14)]
    assert fx.assign_layers(al, "autoencoder", aref) == len(aref)


def test_saliency_front_end_follows_the_reference(cic):
    """compute_saliency_map is the GPU path (no cv2.saliency route): an unknown method raises like GAN_functions.py:110 before any
    device work, and without a device the call fails loudly instead of falling back (create_saliency_mask likewise)."""
    import torch
    sal = cic.saliency
    img = np.zeros((8, 8, 3), np.float32)
    with pytest.raises(ValueError, match="Unsupported"):
        sal.compute_saliency_map(img, method="nope")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            sal.compute_saliency_map(img, method="combined")
        with pytest.raises(RuntimeError):
            sal.create_saliency_mask(np.zeros((8, 8), np.float32), smooth=False)
    assert not hasattr(sal, "_saliency_module")


def test_model_caches_follow_the_plan_generation(cic):
    """set_weights_dict drops the plan AND everything that holds raw pointers into it (graph caches were keyed by id())."""
    import train_autoencoder as tr
    m = tr.build_autoencoder((32, 32, 3))
    m.__dict__["_pipe_graphs"] = {"k": 1}
    m.__dict__["_phase_graphs"] = {"k": 1}
    m.__dict__["_phase_cache"] = {"k": 1}
    assert m._weights is None                                           # Keras-default weights are drawn lazily
    m.set_weights_dict(cic.weights.synthetic_autoencoder(seed=1))
    assert not any(k in m.__dict__ for k in ("_pipe_graphs", "_phase_graphs", "_phase_cache")) and m._plan is None
    shapes = cic.weights.adaptive_shapes((256, 256, 3), 512)            # no allocation: zero-stride views
    assert shapes["hq_encoder"]["dense/kernel"].shape == (131072, 1024) and shapes["hq_encoder"]["dense/kernel"].strides == (0, 0)


def test_import_shim_does_not_duplicate_modules(cic):
    """`cic_b200.x` and `contextual-image-compression_b200.x` are one module object (a second copy would split module-level state
    such as the precision switch or the workspace cache)."""
    import importlib
    import GAN_functions, GAN_test, test_autoencoder, train_autoencoder  # noqa: F401  (the drop-ins import through the shim)
    import cic_b200.models as m2
    import cic_b200.runtime as r2
    assert importlib.import_module("contextual-image-compression_b200.models") is m2
    assert importlib.import_module("contextual-image-compression_b200.runtime") is r2 and cic.runtime is r2 and m2.runtime is r2
    assert GAN_functions.build_encoder.__module__ == "contextual-image-compression_b200.gan"


def test_bench_clock_sampler_falls_back_to_the_nvidia_smi_loop():
    """bench.py's clock block: when the NVML child produced nothing, the samples of the nvidia-smi loop that fall inside the timed
    region are used (median clock, throttle reasons that were Active)."""
    import importlib.util
    import time
    import datetime
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Fake:
        def __init__(self, out):
            self.out = out

        def terminate(self):
            pass

        def kill(self):
            pass

        def communicate(self, timeout=None):
            return self.out, ""

    now = time.time()
    stamp = lambda t: datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]   # noqa: E731
    lines = [f"{stamp(now - 5.0)}, 1965, 1965, Not Active, Not Active, Not Active, Not Active",
             f"{stamp(now + 0.10)}, 1650, 1965, Not Active, Not Active, Not Active, Active",
             f"{stamp(now + 0.20)}, 1700, 1965, Not Active, Not Active, Not Active, Active",
             f"{stamp(now + 0.30)}, 1680, 1965, Not Active, Not Active, Not Active, Not Active",
             "garbage line"]
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.proc, s.err, s.smi = Fake(""), None, Fake("\n".join(lines) + "\n")
    s.t0 = now
    time.sleep(0.45)
    got = s.stop()
    assert got["source"].startswith("nvidia-smi") and got["samples"] == 3
    assert got["sm_mhz"] == 1680.0 and got["sm_max_mhz"] == 1965 and got["reasons"] == ["sw_power_cap"]
    # the NVML child's samples win when it produced any
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.proc, s.err, s.smi = Fake(f"max 1965\n{now + 0.1!r} 1800 sw_power_cap\n"), None, Fake("\n".join(lines) + "\n")
    s.t0 = now
    got = s.stop()
    assert got["source"] == "nvml" and got["sm_mhz"] == 1800.0
