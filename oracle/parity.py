"""Parity report of a GPU run of the tiled adaptive codec against the CPU oracle.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

The three north-star criteria (BASELINE.json), evaluated on the tiles the oracle was run on:
  * quantised latent symbols bit-exact except where the oracle's fp32 pre-round value lies within 1e-3 of a rounding boundary
    (mismatches counted in total and outside that band);
  * reconstruction within 1e-2 max-abs in [0,1] pixel space - on tiles that decoded the same symbols (a tile whose symbol
    flipped inside the band decodes a different latent by construction);
  * PSNR within 0.05 dB (same tiles), computed with the skimage restatement of oracle/metrics.py on the part of the tile that
    lies inside the image.
Used by tests/ (assertions) and by bench.py's `parity` block (reporting).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import metrics, tiling

BAND = 1e-3


def symbol_parity(got_sym, want_sym, want_pre):
    """(#mismatches, #mismatches outside the 1e-3 rounding-boundary band)."""
    bad = np.asarray(got_sym).astype(np.int64) != np.asarray(want_sym).astype(np.int64)
    frac = np.abs(want_pre - np.floor(want_pre))
    near = np.abs(frac - 0.5) < BAND
    return int(bad.sum()), int((bad & ~near).sum())


def adaptive_parity(got: Dict[str, np.ndarray], ref: Dict[str, np.ndarray], n: int, h: int, w: int, tile: int,
                    max_metric_tiles: int = 8) -> Dict[str, float]:
    """`got`: full-batch GPU outputs (blended (n,h,w,3), dt (n,h,w,1), hq_symbols / lq_symbols (n_tiles, L)); `ref`: the result
    of tiling.adaptive_forward_tiled on some of the tiles."""
    sel = ref["tiles"]
    rep: Dict[str, float] = {"tiles_checked": int(len(sel)), "tile": tile}
    same = np.ones(len(sel), bool)
    for br in ("hq", "lq"):
        g = np.asarray(got[f"{br}_symbols"])[sel]
        bad, outside = symbol_parity(g, ref[f"{br}_sym"], ref[f"{br}_pre"])
        rep[f"{br}_symbols"] = int(g.size)
        rep[f"{br}_symbol_mismatches"] = bad
        rep[f"{br}_symbol_mismatches_outside_band"] = outside
        same &= (g.astype(np.int64) == ref[f"{br}_sym"].astype(np.int64)).all(axis=1)
    rep["symbol_mismatches"] = rep["hq_symbol_mismatches"] + rep["lq_symbol_mismatches"]
    rep["symbol_mismatches_outside_band"] = rep["hq_symbol_mismatches_outside_band"] + rep["lq_symbol_mismatches_outside_band"]
    rep["tiles_with_identical_symbols"] = int(same.sum())
    recon, dts, dpsnr, dssim = 0.0, 0.0, 0.0, 0.0
    done = 0
    for k, t in enumerate(sel):
        img, y0, x0, vh, vw = tiling.tile_window(int(t), h, w, tile)
        gb = np.asarray(got["blended"])[img, y0:y0 + vh, x0:x0 + vw]
        wb = ref["blended"][k, :vh, :vw]
        dts = max(dts, float(np.abs(np.asarray(got["dt"])[img, y0:y0 + vh, x0:x0 + vw] - ref["dt"][k, :vh, :vw]).max()))
        if not same[k]:
            continue
        recon = max(recon, float(np.abs(gb - wb).max()) / 2.0)
        if done < max_metric_tiles and vh >= 7 and vw >= 7:
            orig = ref["img_tiles"][k, :vh, :vw]
            a, b = metrics.compute_metrics(orig, wb), metrics.compute_metrics(orig, gb)
            dpsnr = max(dpsnr, abs(a["psnr"] - b["psnr"]))
            dssim = max(dssim, abs(a["ssim"] - b["ssim"]))
            done += 1
    rep["recon_max_abs_01"] = recon
    rep["dt_max_abs"] = dts
    rep["psnr_delta_db_max"] = dpsnr
    rep["ssim_delta_max"] = dssim
    rep["metric_tiles"] = done
    return rep


def assert_north_star(rep: Dict[str, float], recon_tol: float = 1e-2, psnr_tol: float = 0.05) -> None:
    assert rep["symbol_mismatches_outside_band"] == 0, rep
    assert rep["tiles_with_identical_symbols"] > 0, rep
    assert rep["recon_max_abs_01"] < recon_tol, rep
    assert rep["psnr_delta_db_max"] < psnr_tol, rep
    assert rep["dt_max_abs"] < 1e-5, rep
