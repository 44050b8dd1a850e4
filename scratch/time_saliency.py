"""Device time of the saliency front end (compute_saliency_map 'combined' -> create_saliency_mask) against the CPU restatement built
on the real OpenCV core routines; prints one JSON line.  Usage: python scratch/time_saliency.py > gpurun_out/r02_saliency.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cic_b200 as cic  # noqa: E402
from oracle import saliency as osal  # noqa: E402
from test_oracle_saliency import photo  # noqa: E402

out = {}
for name, b, h, w in (("64x256x256", 64, 256, 256), ("64x512x512", 64, 512, 512), ("8x1080p", 8, 1080, 1920)):
    bgr = np.stack([photo(h, w, seed=i) for i in range(min(b, 4))])
    bgr = np.concatenate([bgr] * (b // len(bgr)))
    rgb = torch.from_numpy(np.ascontiguousarray(bgr[..., ::-1])).cuda()
    row = {}
    for label, fn in (("map_combined", lambda: cic.ops.saliency_map(rgb, "combined")),
                      ("map_spectral", lambda: cic.ops.saliency_map(rgb, "spectral_residual")),
                      ("map_fine", lambda: cic.ops.saliency_map(rgb, "fine_grained")),
                      ("map_and_mask", lambda: cic.ops.saliency_mask_from_image(rgb, "combined"))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        row[label + "_ms"] = e0.elapsed_time(e1) / 10
    t = time.perf_counter()
    n_cpu = 2
    for i in range(n_cpu):
        m = osal.compute_saliency_map(np.ascontiguousarray(bgr[i][..., ::-1]), "combined", use_cv=True)
        osal.create_saliency_mask(m, smooth=True)
    row["cpu_ms_per_image"] = (time.perf_counter() - t) / n_cpu * 1e3
    row["gpu_mpix_s"] = b * h * w / row["map_and_mask_ms"] / 1e3
    row["cpu_mpix_s"] = h * w / row["cpu_ms_per_image"] / 1e3
    out[name] = row
print(json.dumps(out))
