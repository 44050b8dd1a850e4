"""Probe: does a UMMA smem descriptor whose start address is shifted by s rows (s*128 B, not a multiple of the
1024-B swizzle repeat) read the rows TMA wrote, with base_offset = 0 or = s?  1x1 conv, W = 128, TH = 1."""
import os, sys, subprocess, json
import numpy as np
if len(sys.argv) > 1:
    s, bo, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    os.environ["CIC_TC_DBG_SHIFT"] = str(s); os.environ["CIC_TC_DBG_BO"] = str(bo)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch, cic_b200 as cic
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 2, 128, C)).astype(np.float32)
    k = (rng.standard_normal((1, 1, C, 32)) / np.sqrt(C)).astype(np.float32)
    got = cic.ops.conv2d_tc(x, k).cpu().numpy()
    xr = torch.from_numpy(x).bfloat16().float().numpy(); kr = torch.from_numpy(k).bfloat16().float().numpy()
    want = xr.reshape(-1, C) @ kr.reshape(C, 32)
    err = np.abs(got.reshape(-1, 32) - want).max(axis=1).reshape(2, 128)
    ok = err < 1e-3
    print(json.dumps({"shift": s, "bo": bo, "C": C, "rows_ok_row0": int(ok[0].sum()), "first_bad": int(np.argmin(ok[0])) if not ok[0].all() else -1,
                      "ok_prefix_len": int(np.argmin(ok[0])) if not ok[0].all() else 128, "pattern": "".join("1" if v else "0" for v in ok[0][:32])}))
else:
    for C in (64, 32):
        for s in (0, 1, 2, 3, 5, 8):
            for bo in sorted({0, s % 8}):
                r = subprocess.run([sys.executable, __file__, str(s), str(bo), str(C)], capture_output=True, text=True)
                print(r.stdout.strip() or r.stderr[-300:])
