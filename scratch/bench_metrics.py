"""Times cic_metrics_psnr_ssim_f32_fast at the C2 shape (64 x 512x512 RGB) for several strip heights (tuning build only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cic_b200 as cic
n, h, w = 64, 512, 512
a = torch.rand((n, h, w, 3), device="cuda") * 2 - 1
b = (a + 0.05 * torch.randn_like(a)).clamp(-1, 1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rows in [0, 64, 128] if len(sys.argv) < 2 else [int(sys.argv[1])]:
    os.environ["CIC_METRICS_ROWS"] = str(rows)
    for strip, packed in (("1", "1"), ("1", "0"), ("0", "0")):
        os.environ["CIC_METRICS_STRIP"] = strip
        os.environ["CIC_METRICS_PACKED"] = packed
        cic.ops.metrics_f32(a, b, signed_range=True, fast=True)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(20):
            m = cic.ops.metrics_f32(a, b, signed_range=True, fast=True)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 20
        print(f"rows={rows} strip={strip} packed={packed}: {ms:.4f} ms  {24.0 * n * h * w / ms / 1e6:.0f} GB/s  ssim {m[:, 1].mean().item():.9f} psnr {m[:, 0].mean().item():.9f}")
