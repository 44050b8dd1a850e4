"""Prints the stream time line of predict_stream at the C2 shape (debug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cic_b200 as cic
import GAN_functions as gf
u8 = len(sys.argv) < 2 or sys.argv[1] != "f32"
models = gf.build_adaptive_compression_model((256, 256, 3), 512, target_bpp=True)
am = models["adaptive_model"]
am.set_weights_dict(cic.weights.synthetic_adaptive((256, 256, 3), 512, seed=42))
n = 64
img8 = cic.synth.synth_images_u8(n, 512, 512, seed=43)
src = [torch.from_numpy(img8 if u8 else cic.synth.to_signed_range(img8)).pin_memory(), torch.from_numpy(cic.synth.synth_masks(n, 512, 512, seed=43)).pin_memory(),
       torch.ones((n, 1)).pin_memory()]
def on_batch(d_in, outs):
    m = cic.ops.metrics_f32(d_in[0], outs["blended"], signed_range=True, fast=True)
    return cic.ops.metric_sums(m, outs["hq_ratio_sum"], 512 * 512, 1024, 512, 65536)
for _ in am.predict_stream((src for _ in range(4)), on_batch=on_batch, u8_io=u8):
    pass
cic.runtime.set_pipe_timeline(True)
print("timeline flag", cic.runtime.pipe_timeline(), flush=True)
cnt = 0
for _ in am.predict_stream((src for _ in range(8)), on_batch=on_batch, u8_io=u8):
    cnt += 1
print("batches", cnt, "keys", [k for k in am.__dict__ if k.startswith("_stream")], flush=True)
