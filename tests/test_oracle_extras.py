"""CPU tests of the oracle pieces added in round 2: the tiled-codec rule, the parity report, the rANS restatement, MS-SSIM."""
import numpy as np
import pytest

from oracle import metrics, parity, rans, tiling


def test_tiling_round_trip_and_windows():
    rng = np.random.default_rng(0)
    for (n, h, w, t) in ((2, 5, 7, 4), (1, 8, 8, 4), (3, 3, 10, 4), (1, 1080, 1920, 256)):
        a = rng.random((n, h, w, 2)).astype(np.float32)
        tiles = tiling.split_tiles(a, t)
        ty, tx = tiling.grid(h, w, t)
        assert tiles.shape == (n * ty * tx, t, t, 2)
        np.testing.assert_array_equal(tiling.join_tiles(tiles, n, h, w, t), a)
        # padding replicates the last row / column (numpy mode='edge')
        last = tiles[n * ty * tx - 1]
        img, y0, x0, vh, vw = tiling.tile_window(n * ty * tx - 1, h, w, t)
        assert (img, y0 + vh, x0 + vw) == (n - 1, h, w)
        np.testing.assert_array_equal(last[vh:, :vw], np.broadcast_to(last[vh - 1:vh, :vw], last[vh:, :vw].shape))
        np.testing.assert_array_equal(last[:, vw:], np.broadcast_to(last[:, vw - 1:vw], last[:, vw:].shape))
    assert tiling.grid(1080, 1920, 256) == (5, 8) and tiling.grid(2160, 3840, 256) == (9, 15)      # 40 and 135 tiles
    s = tiling.sample_tiles(135, 12, seed=3)
    assert len(s) == 12 and s[0] == 0 and s[-1] == 134 and len(set(s.tolist())) == 12
    np.testing.assert_array_equal(tiling.sample_tiles(5, 10), np.arange(5))


def test_symbol_band_rule():
    want_pre = np.array([0.4, 0.4995, 0.5005, 1.2, -2.5004, 3.0])
    want_sym = np.rint(want_pre)
    got = want_sym.copy()
    got[1] += 1          # inside the 1e-3 band around .5: tolerated
    got[4] -= 1          # inside the band
    got[3] += 1          # outside: a real mismatch
    assert parity.symbol_parity(got, want_sym, want_pre) == (3, 1)
    assert parity.symbol_parity(want_sym, want_sym, want_pre) == (0, 0)


@pytest.mark.parametrize("shape,scale", [((4, 1024), 3.0), ((3, 70), 20.0), ((2, 512), 0.0), ((1, 32), 400.0), ((0, 64), 1.0), ((2, 33), 5.0)])
def test_rans_restatement_round_trip(shape, scale):
    rng = np.random.default_rng(shape[1])
    x = np.rint(rng.standard_normal(shape) * scale).astype(np.int32)
    stream = rans.encode(x)
    back = rans.decode(stream)
    np.testing.assert_array_equal(back, np.clip(x, -rans.SYM_MAX, rans.SYM_MAX))
    f = rans.build_freq(x)
    assert f.sum() == rans.M and np.all(f[np.bincount((np.clip(x, -1023, 1023) + 1023).ravel(), minlength=rans.ALPHA) > 0] >= 1)
    header = np.frombuffer(stream[:32], "<u4")
    assert header[0] == rans.MAGIC and (header[2], header[3]) == shape
    if x.size > 2000 and scale > 0:                                   # coded size ~ entropy + 32 states per row
        p = np.bincount((x + 1023).ravel(), minlength=rans.ALPHA) / x.size
        ent = -(p[p > 0] * np.log2(p[p > 0])).sum() * x.size
        payload_bits = 8 * (len(stream) - 32 - 4096 - 4 * (shape[0] + 1))
        assert ent <= payload_bits <= ent * 1.03 + 32 * 32 * shape[0] + 64 * shape[0]


def test_rans_streams_decode_row_by_row_independently():
    """Rows are independent streams (tiles stay independently decodable): decoding a container rebuilt from a subset of the rows
    gives those rows."""
    rng = np.random.default_rng(5)
    x = np.rint(rng.laplace(0, 3, (6, 256))).astype(np.int32)
    s = rans.encode(x)
    rows, L = 6, 256
    off0 = 32 + 4096
    offs = np.frombuffer(s[off0:off0 + 4 * (rows + 1)], "<u4").astype(int)
    payload = off0 + 4 * (rows + 1)
    keep = [4, 1]
    blobs = [s[payload + offs[r]:payload + offs[r + 1]] for r in keep]
    hdr = np.array([rans.MAGIC, 1, len(keep), L, rans.PROB_BITS, rans.ALPHA, 0, 0], "<u4").tobytes()
    new_offs = np.concatenate([[0], np.cumsum([len(b) for b in blobs])]).astype("<u4").tobytes()
    sub = hdr + s[32:32 + 4096] + new_offs + b"".join(blobs)
    np.testing.assert_array_equal(rans.decode(sub), x[keep])


def test_ms_ssim_identities():
    rng = np.random.default_rng(1)
    a = rng.random((192, 208, 3))
    assert metrics.ms_ssim(a, a) == pytest.approx(1.0, abs=1e-12)
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1)
    c = np.clip(a + 0.3 * rng.standard_normal(a.shape), 0, 1)
    assert 0 < metrics.ms_ssim(a, c) < metrics.ms_ssim(a, b) < 1                 # more noise, lower score
    assert metrics.ms_ssim(a, b) == pytest.approx(metrics.ms_ssim(b, a), rel=1e-12)   # symmetric
    g = metrics.gaussian_window()
    assert g.sum() == pytest.approx(1.0) and g.argmax() == 5 and np.allclose(g, g[::-1])
    gray = metrics.ms_ssim(a[..., 0], b[..., 0])                                   # 2-D input = one channel
    assert gray == pytest.approx(metrics.ms_ssim(a[..., :1], b[..., :1]))


def test_hq_ratio_of_the_unpadded_image_matches_appendix_c():
    """dt known answers of SURVEY App. C through the tiled-codec helper (mean over the image's own pixels)."""
    mask = np.full((1, 10, 20, 1), 0.75, np.float32)
    for bpp, want in ((0.1, 0.212834), (1.0, 0.852215), (2.0, 0.994246)):
        assert tiling.hq_ratio(mask, np.array([bpp]))[0] == pytest.approx(want, abs=2e-6)


def _jpeg_test_image(h, w, kind, seed):
    import cv2
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return (rng.random((h, w, 3)) * 255).astype(np.uint8)
    if kind == "flat":
        return np.full((h, w, 3), (7, 250, 128), np.uint8)
    if kind == "saturated":                                  # long runs of 0 / 255: exercises 0xFF stuffing and ZRL
        return (((np.indices((h, w)).sum(0) // 3) % 2) * 255).astype(np.uint8)[..., None].repeat(3, axis=2)
    base = rng.random((h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC) + 0.03 * rng.standard_normal((h, w, 3))
    return np.clip(img * 255, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w", [(16, 16), (64, 64), (48, 80), (100, 150), (17, 33), (8, 8), (1, 1), (120, 68), (99, 151), (128, 128)])
def test_jpeg_restatement_is_byte_identical_to_opencv(h, w):
    """oracle/jpeg.py against the real library behind the reference's cv2.imwrite (test_autoencoder.py:93, GAN_functions.py:50)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import jpeg
    for kind in ("noise", "smooth", "flat", "saturated"):
        img = _jpeg_test_image(h, w, kind, h * 1000 + w)
        assert jpeg.encode_bgr(img) == bytes(cv2.imencode(".jpg", img)[1]), (kind, "default quality")
    img = _jpeg_test_image(h, w, "smooth", 5)
    for q in (100, 75, 50, 10):
        assert jpeg.encode_bgr(img, q) == bytes(cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])[1]), q


def test_jpeg_restatement_decodes():
    cv2 = pytest.importorskip("cv2")
    from oracle import jpeg
    img = _jpeg_test_image(96, 112, "smooth", 9)
    back = cv2.imdecode(np.frombuffer(jpeg.encode_bgr(img), np.uint8), cv2.IMREAD_COLOR)
    assert back.shape == img.shape
    mse = np.mean((back.astype(np.float64) - img) ** 2)
    assert 10 * np.log10(255 ** 2 / mse) > 28            # q95 with 4:2:0 chroma on an image carrying 3 % noise
    with pytest.raises(ValueError):
        jpeg.encode_bgr(img[..., 0])
