"""CPU oracle of the TILED codec.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference graph is fixed at 256x256 (GAN_functions.py:242-248: literal skip shapes, Dense(16*16*512) :247, Flatten -> Dense
:325-326; IMG_SIZE = (256, 256) GAN_test.py:23).  BASELINE.json's 512^2 / 1024^2 / 1080p / 4K configs are therefore coded as
independent 256x256 tiles: the image and its saliency mask are extended to multiples of the tile by
replicating the last row / column (numpy `mode='edge'`), every tile goes through the reference graph with its image's target
bpp, outputs are cropped back, and hq_ratio is the mean of dt over the pixels of the unpadded image.  This file restates that
rule in numpy around `graphs.adaptive_forward`; parity is per tile against the oracle on identical tiles.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import graphs


def grid(h: int, w: int, tile: int):
    """(tiles down, tiles across) of an h x w image."""
    return -(-h // tile), -(-w // tile)


def pad_to_tiles(a: np.ndarray, tile: int) -> np.ndarray:
    """(n,H,W,C) -> (n, ceil(H/tile)*tile, ceil(W/tile)*tile, C), edge-replicated."""
    n, h, w, c = a.shape
    ty, tx = grid(h, w, tile)
    if ty * tile == h and tx * tile == w:
        return a
    return np.pad(a, ((0, 0), (0, ty * tile - h), (0, tx * tile - w), (0, 0)), mode="edge")


def split_tiles(a: np.ndarray, tile: int) -> np.ndarray:
    """(n,H,W,C) -> (n*ty*tx, tile, tile, C) in (image, tile row, tile column) order, after edge padding."""
    p = pad_to_tiles(a, tile)
    n, hp, wp, c = p.shape
    ty, tx = hp // tile, wp // tile
    return p.reshape(n, ty, tile, tx, tile, c).transpose(0, 1, 3, 2, 4, 5).reshape(n * ty * tx, tile, tile, c)


def join_tiles(t: np.ndarray, n: int, h: int, w: int, tile: int) -> np.ndarray:
    """Inverse of split_tiles, cropped to the unpadded h x w."""
    ty, tx = grid(h, w, tile)
    c = t.shape[-1]
    full = t.reshape(n, ty, tx, tile, tile, c).transpose(0, 1, 3, 2, 4, 5).reshape(n, ty * tile, tx * tile, c)
    return full[:, :h, :w]


def tile_window(index: int, h: int, w: int, tile: int):
    """Global tile index -> (image, y0, x0, valid rows, valid columns): the part of the tile that lies inside the image."""
    ty, tx = grid(h, w, tile)
    img, r = divmod(index, ty * tx)
    y0, x0 = (r // tx) * tile, (r % tx) * tile
    return img, y0, x0, min(tile, h - y0), min(tile, w - x0)


def sample_tiles(n_tiles: int, count: int, seed: int = 0) -> np.ndarray:
    """`count` distinct tile indices spread over the batch (always including the first and the last tile: the last one is the
    ragged corner of the last image)."""
    if count >= n_tiles:
        return np.arange(n_tiles)
    rng = np.random.Generator(np.random.PCG64(seed))
    mid = rng.choice(np.arange(1, n_tiles - 1), size=count - 2, replace=False)
    return np.sort(np.concatenate([[0, n_tiles - 1], mid]))


def adaptive_forward_tiled(weights, img: np.ndarray, mask: np.ndarray, bpp: np.ndarray, tile: int,
                           tiles: Optional[Sequence[int]] = None, chunk: int = 16) -> Dict[str, np.ndarray]:
    """The adaptive model on the (sampled) tiles of a batch of images of any size.

    Returns {'tiles': global tile indices, 'blended' (k,tile,tile,3), 'dt' (k,tile,tile,1), 'hq_latent_q', 'lq_latent_q',
    'rd_params', and the extras of graphs.adaptive_forward (hq_sym, hq_pre, hq_out, ...)} for the k selected tiles, run through the
    oracle `chunk` tiles at a time."""
    n, h, w, _ = img.shape
    ty, tx = grid(h, w, tile)
    ti, tm = split_tiles(img, tile), split_tiles(mask, tile)
    tb = np.repeat(np.asarray(bpp, np.float32).reshape(-1), ty * tx).reshape(-1, 1)
    sel = np.arange(n * ty * tx) if tiles is None else np.asarray(tiles, dtype=np.int64)
    acc: Dict[str, list] = {}
    names = ["blended", "hq_latent_q", "lq_latent_q", "rd_params", "dt"]
    for i in range(0, len(sel), chunk):
        s = sel[i:i + chunk]
        outs, ex = graphs.adaptive_forward(weights, np.ascontiguousarray(ti[s]), np.ascontiguousarray(tm[s]), tb[s], return_extras=True)
        for k, v in zip(names, outs):
            acc.setdefault(k, []).append(v)
        for k, v in ex.items():
            if isinstance(v, (list, tuple)):
                continue
            acc.setdefault(k, []).append(np.asarray(v))
    res = {k: np.concatenate(v, axis=0) for k, v in acc.items()}
    res["tiles"] = sel
    res["img_tiles"] = ti[sel]
    return res


def hq_ratio(mask: np.ndarray, bpp: np.ndarray) -> np.ndarray:
    """mean(dt) over the unpadded pixels of every image (GAN_test.py:312), float64."""
    dt = graphs.dynamic_threshold(mask, np.asarray(bpp, np.float32).reshape(-1, 1))
    return dt.reshape(dt.shape[0], -1).mean(axis=1, dtype=np.float64)
