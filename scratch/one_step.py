"""One C2 bench step (64 x 512x512 = 256 tiles: forward + PSNR/SSIM + metric sums) between cudaProfilerStart/Stop, after two warm steps:
  ncu --profile-from-start off ... python scratch/one_step.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cic_b200 as cic
import GAN_functions as gf
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
models = gf.build_adaptive_compression_model((256, 256, 3), 512, target_bpp=True)
am = models["adaptive_model"]
am.set_weights_dict(cic.weights.synthetic_adaptive((256, 256, 3), 512, seed=42))
n = 64
img = torch.from_numpy(cic.synth.to_signed_range(cic.synth.synth_images_u8(n, 512, 512, seed=43))).cuda()
mask = torch.from_numpy(cic.synth.synth_masks(n, 512, 512, seed=43)).cuda()
bpp = torch.ones((n,), device="cuda")
def step():
    am.forward_device([img, mask, bpp])
    m = cic.ops.metrics_f32(img, am.last["blended"], signed_range=True, fast=True)
    return cic.ops.metric_sums(m, am.last["hq_ratio_sum"], 512 * 512, 1024, 512, 65536)
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
