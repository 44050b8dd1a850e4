"""Times the RD optimizer plan on 256 tiles for the NPX variants of its first conv (tuning build only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cic_b200 as cic
import GAN_functions as gf
rd = gf.build_rate_distortion_optimizer((256, 256, 3), None)
rd.set_weights_dict(cic.weights.synthetic_rd_optimizer(seed=4))
mask = torch.from_numpy(cic.synth.synth_masks(8, 256, 256, seed=4)).cuda().repeat(32, 1, 1, 1).contiguous()
bpp = torch.ones((256, 1), device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ref = None
for npx in ("2", "1"):
    os.environ["CIC_RD_NPX"] = npx
    out = rd.forward_device([mask, bpp])[0]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(20):
        out = rd.forward_device([mask, bpp])[0]
    ev[1].record()
    torch.cuda.synchronize()
    if ref is None: ref = out.clone()
    print(f"NPX={npx}: rd plan {ev[0].elapsed_time(ev[1]) / 20:.4f} ms; max diff vs NPX=2 {float((out - ref).abs().max()):.2e}")
