"""The package's own HDF5 reader and the Keras legacy-H5 layer on top of it (GAN_test.py:37-78).

Two kinds of evidence:
  * `test_real_libhdf5_file`: a file written by a real libhdf5 (MATLAB 7.4, shipped with scipy's test data) - the reader's parsing of
    the superblock, user block, symbol-table group, version-1 object header, dataspace / datatype / layout messages and a string
    attribute is checked against the same variable read by scipy from the v5 MAT file next to it;
  * the rest: files written by tests/h5_writer.py in the Keras layout, read back through keras_h5 / GAN_test.load_models.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import h5_writer  # noqa: E402

import cic_b200  # noqa: E402
from cic_b200 import hdf5_lite, keras_h5, weights  # noqa: E402


def _scipy_data():
    try:
        import scipy.io.matlab
    except ImportError:
        return None
    d = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data")
    return d if os.path.exists(os.path.join(d, "testhdf5_7.4_GLNX86.mat")) else None


@pytest.mark.skipif(_scipy_data() is None, reason="scipy's MATLAB test data is not installed")
def test_real_libhdf5_file():
    import scipy.io
    d = _scipy_data()
    f = hdf5_lite.File(os.path.join(d, "testhdf5_7.4_GLNX86.mat"))          # HDF5 behind MATLAB's 512-byte user block
    assert f.r.base == 512 and f.keys() == ["testdouble"]
    ds = f["testdouble"]
    assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8")
    assert bytes(ds.attrs["MATLAB_class"]) == b"double"
    want = scipy.io.loadmat(os.path.join(d, "testdouble_7.4_GLNX86.mat"))["testdouble"]   # the same variable, v5 MAT format
    np.testing.assert_array_equal(ds.read().T, want)                         # MATLAB stores column-major
    with pytest.raises(KeyError):
        f["missing"]


def test_not_hdf5():
    with pytest.raises(hdf5_lite.Hdf5Error, match="not an HDF5 file"):
        hdf5_lite.File(b"PK\x03\x04" + b"\x00" * 4096)


@pytest.mark.parametrize("split,user_block", [(False, 0), (True, 0), (True, 512)])
def test_writer_reader_round_trip(split, user_block):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 4, 5)).astype(np.float32)
    b = rng.integers(-5, 5, (7,)).astype(np.int32)
    many = {f"d{i:02d}": np.full((2,), i, np.float64) for i in range(21)}     # three symbol-table nodes
    data = h5_writer.write_tree({"g": ({"a": (a, {"note": np.bytes_(b"hello"), "n": np.int32(7)}), "b": b}, {"names": np.array([b"x", b"yy"], "S2")}),
                                 "many": many, "empty": np.zeros((0, 3), np.float32)},
                                attrs={"top": np.float32(1.5), "text": h5_writer.VlenStr("vari\u00e9 length")}, user_block=user_block, split=split)
    f = hdf5_lite.File(data)
    assert sorted(f.keys()) == ["empty", "g", "many"] and f.attrs["top"] == np.float32(1.5)
    assert f.attrs["text"].decode("utf-8") == "vari\u00e9 length"               # variable-length string through the global heap
    np.testing.assert_array_equal(f["g/a"].read(), a)
    np.testing.assert_array_equal(f["g"]["b"].read(), b)
    assert f["g/a"].attrs["note"] == b"hello" and f["g/a"].attrs["n"] == 7
    assert list(f["g"].attrs["names"]) == [b"x", b"yy"]
    assert f["many"].keys() == sorted(many)
    for k, v in many.items():
        np.testing.assert_array_equal(f["many"][k].read(), v)
    assert f["empty"].read().shape == (0, 3)


@pytest.mark.parametrize("deflate,shuffle", [(False, False), (True, False), (True, True)])
def test_chunked_datasets(deflate, shuffle):
    # Keras writes contiguous datasets; h5py users who re-save with compression get chunked + shuffle + deflate
    rng = np.random.default_rng(1)
    a = rng.standard_normal((10, 7, 3)).astype(np.float32)
    data = h5_writer.write_tree({"a": h5_writer.Chunked(a, (4, 4, 3), deflate, shuffle)})
    np.testing.assert_array_equal(hdf5_lite.File(data)["a"].read(), a)


def _keras_layers(nested, sub, kind):
    """(class, layer name, [(weight name, array)]) in the order Keras lists the layers of the reference's models, from our flat names"""
    ws = nested[sub]
    out, counters = [("InputLayer", "input_1", [])], {}

    def nm(base):
        i = counters.get(base, 0)
        counters[base] = i + 1
        return base if i == 0 else f"{base}_{i}"

    def conv(key, cls="Conv2D", base="conv2d"):
        n = nm(base)
        out.append((cls, n, [(f"{n}/kernel:0", ws[key + "/kernel"]), (f"{n}/bias:0", ws[key + "/bias"])]))

    def bn(key):
        n = nm("batch_normalization")
        out.append(("BatchNormalization", n, [(f"{n}/{t}:0", ws[f"{key}/{t}"]) for t in keras_h5.BN_NAMES]))

    def dense(key):
        conv(key, "Dense", "dense")

    if kind == "encoder":
        conv("conv1"); out.append(("LeakyReLU", "leaky_re_lu", []))
        conv("conv2"); bn("bn2"); conv("conv3"); bn("bn3")
        if "attn/gamma" in ws:
            n = "self_attention"
            sub_names = [nm("conv2d") for _ in range(3)]
            wl = [(f"{n}/gamma:0", ws["attn/gamma"])]
            for s, k in zip(sub_names, ("query", "key", "value")):
                wl += [(f"{n}/{s}/kernel:0", ws[f"attn/{k}/kernel"]), (f"{n}/{s}/bias:0", ws[f"attn/{k}/bias"])]
            out.append(("SelfAttention", n, wl))
        conv("conv4"); bn("bn4"); out.append(("Flatten", "flatten", [])); dense("dense")
    elif kind == "generator":
        dense("dense"); bn("bn0"); out.append(("Reshape", "reshape", []))
        for i in range(1, 5):
            conv(f"deconv{i}", "Conv2DTranspose", "conv2d_transpose"); bn(f"bn{i}")
            out.append(("Concatenate", nm("concatenate"), []))
        conv("conv_out")
    elif kind == "latent_saliency":
        for i in (1, 2, 3):
            dense(f"dense{i}")
    elif kind == "rd_optimizer":
        conv("conv1"); conv("conv2"); out.append(("GlobalAveragePooling2D", "gap", [])); dense("dense1"); dense("dense2")
    return out


IMG, LAT = (32, 32, 3), 8


@pytest.fixture(scope="module")
def small_adaptive():
    return weights.synthetic_adaptive(IMG, LAT, seed=3)


def test_keras_component_files(tmp_path, small_adaptive):
    for i, (sub, kind) in enumerate(keras_h5.SUB_MODELS):
        data = h5_writer.write_keras(_keras_layers(small_adaptive, sub, kind), model_name=sub, split=bool(i % 2))
        (tmp_path / f"{sub}_final.h5").write_bytes(data)
    got = keras_h5.load_adaptive_dir(str(tmp_path))
    weights.check_adaptive(got, IMG, LAT)
    for sub, ws in small_adaptive.items():
        assert set(got[sub]) == set(ws)
        for k, v in ws.items():
            np.testing.assert_array_equal(got[sub][k], v, err_msg=f"{sub}/{k}")


def test_keras_latest_epoch_and_missing(tmp_path, small_adaptive):
    assert keras_h5.find_suffix(str(tmp_path)) is None
    with pytest.raises(FileNotFoundError):
        keras_h5.load_adaptive_dir(str(tmp_path))
    for epoch in (10, 20):
        for sub, kind in keras_h5.SUB_MODELS:
            ws = {s: {k: v + np.float32(epoch) for k, v in w.items()} for s, w in small_adaptive.items()}
            (tmp_path / f"{sub}_epoch_{epoch}.h5").write_bytes(h5_writer.write_keras(_keras_layers(ws, sub, kind), model_name=sub))
    assert keras_h5.find_suffix(str(tmp_path)) == "_epoch_20.h5"                       # GAN_test.py:84-97
    got = keras_h5.load_adaptive_dir(str(tmp_path))
    np.testing.assert_array_equal(got["hq_encoder"]["conv1/bias"], small_adaptive["hq_encoder"]["conv1/bias"] + np.float32(20))


def test_keras_weights_only_file_without_config(small_adaptive):
    # model.save_weights("x.h5"): layer groups at the root, no model_config -> classes from Keras' default layer names / shapes
    data = h5_writer.write_keras(_keras_layers(small_adaptive, "hq_generator", "generator"), weights_only=True)
    got = keras_h5.map_layers(keras_h5.read_layers(data), "generator")
    for k, v in small_adaptive["hq_generator"].items():
        np.testing.assert_array_equal(got[k], v, err_msg=k)


def test_keras_layer_order_independent_of_listing(small_adaptive):
    # the counter in Keras' default names is the creation order even if the file lists the layers differently
    layers = _keras_layers(small_adaptive, "rd_optimizer", "rd_optimizer")
    data = h5_writer.write_keras(layers[::-1])
    got = keras_h5.map_layers(keras_h5.read_layers(data), "rd_optimizer")
    for k, v in small_adaptive["rd_optimizer"].items():
        np.testing.assert_array_equal(got[k], v, err_msg=k)


def test_keras_autoencoder_file(tmp_path):
    ws = weights.synthetic_autoencoder(seed=5)
    names = ("conv1", "conv2", "conv3", "conv_x2", "conv5", "conv_x1", "conv_out")
    layers = [("InputLayer", "input_1", [])]
    for i, n in enumerate(names):
        ln = "conv2d" if i == 0 else f"conv2d_{i}"
        layers.append(("Conv2D", ln, [(f"{ln}/kernel:0", ws[n + "/kernel"]), (f"{ln}/bias:0", ws[n + "/bias"])]))
        if i in (0, 1):
            layers.append(("MaxPooling2D", f"max_pooling2d_{i}", []))
    p = tmp_path / "autoencoder_model.h5"
    p.write_bytes(h5_writer.write_keras(layers))
    got = keras_h5.load_autoencoder(str(p))
    assert set(got) == set(ws)
    for k, v in ws.items():
        np.testing.assert_array_equal(got[k], v, err_msg=k)


def test_keras_errors(small_adaptive):
    with pytest.raises(ValueError, match="layer_names"):
        keras_h5.read_layers(h5_writer.write_tree({"x": np.zeros(3, np.float32)}))
    layers = [l for l in _keras_layers(small_adaptive, "hq_encoder", "encoder") if l[0] != "Dense"]
    with pytest.raises(ValueError, match="unexpected layer counts"):
        keras_h5.map_layers(keras_h5.read_layers(h5_writer.write_keras(layers)), "encoder")


def test_load_models_reads_h5_dir_shape_check(tmp_path, small_adaptive):
    # GAN_test.load_models(model_dir) goes through the .h5 route and checks the tensors against the full-size architecture: the small
    # test checkpoint must be rejected by name and shape, an empty directory with the reference's message (GAN_test.py:219)
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import GAN_test
    with pytest.raises(ValueError, match="No models found"):
        GAN_test.load_models(str(tmp_path))
    for sub, kind in keras_h5.SUB_MODELS:
        (tmp_path / f"{sub}_final.h5").write_bytes(h5_writer.write_keras(_keras_layers(small_adaptive, sub, kind), model_name=sub))
    with pytest.raises(ValueError, match=r"hq_encoder/dense/kernel.*has shape"):
        GAN_test.load_models(str(tmp_path))


@pytest.mark.gpu
def test_full_size_h5_directory_drives_the_codec(tmp_path):
    """The reference's own flow at its own sizes: seven `<component>_final.h5` files -> GAN_test.load_models(dir) ->
    compress_and_reconstruct; the result equals the one of a model given the same tensors directly."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import GAN_test
    from cic_b200 import gan, synth
    w = weights.synthetic_adaptive(gan.IMG_SHAPE, gan.BASE_LATENT_DIM, seed=11)
    for i, (sub, kind) in enumerate(keras_h5.SUB_MODELS):
        (tmp_path / f"{sub}_final.h5").write_bytes(h5_writer.write_keras(_keras_layers(w, sub, kind), model_name=sub, split=bool(i % 2)))
    from_files = GAN_test.load_models(str(tmp_path))
    direct = GAN_test.load_models(weights=w)
    img = synth.to_signed_range(synth.synth_images_u8(1, *gan.IMG_SIZE, seed=5))[0]
    mask = synth.synth_masks(1, *gan.IMG_SIZE, seed=6)[0, ..., 0]
    a = GAN_test.compress_and_reconstruct(img, from_files, target_bpp=1.0, mask=mask)
    b = GAN_test.compress_and_reconstruct(img, direct, target_bpp=1.0, mask=mask)
    np.testing.assert_array_equal(a["compressed_img"], b["compressed_img"])
    np.testing.assert_array_equal(a["hq_latent"], b["hq_latent"])
    assert a["actual_bpp"] == b["actual_bpp"]
    assert a["metrics"]["psnr"] == pytest.approx(b["metrics"]["psnr"], rel=1e-12)      # the metric sums are atomics: order-dependent last bits
