#!/usr/bin/env python
"""Headline benchmark: encode + quantise + decode MPix/s of the contextual (adaptive GAN) codec.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision tc|fp32]

Workload (BASELINE.json configs[1]): batch of 64 synthetic 512x512 RGB images per GPU, coded as 256
tiles of 256x256 by the adaptive model of build_adaptive_compression_model (HQ + LQ encoder, latent
saliency, quantiser, two generators, ROI blend) followed by the PSNR/SSIM/bpp evaluation.  One "step" is
one pass over that batch.  N > 1 (torchrun, one rank per GPU): every rank codes its own 64 images
(weak scaling, no data-path collective); the only exchange is the all-reduce of the metric sums, inside
the timed region.  One JSON line is printed by rank 0.

`value`  : MPix/s with inputs resident in HBM, timed with CUDA events on the launch stream.
`e2e`    : same metric through the public API (adaptive_model.predict_phased, or predict_pipelined with --e2e-mode pipelined)
           with pinned HOST buffers, i.e. host->device copy of images/masks/bpp and device->host read of every model output
           per step, overlapped with the kernels chunk by chunk.
`roofline`: the conv/dense GEMM kernels (dominant), algorithmic FLOPs / summed per-layer device time
           (CUDA events recorded around every layer inside the timed steps) against the measured bf16 peak.
`cpu_baseline`: the CPU oracle (torch fp32 restatement of the reference graph) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

IMG_HW = 512
IMGS_PER_GPU = 64
TILE = 256
BASE_LATENT = 512
TARGET_BPP = 1.0
FLOP_PER_TILE = 23.818e9          # algorithmic FLOPs of the adaptive model per 256x256 tile
METRIC = "encode+quantize+decode MPix/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
while True:
    t = time.time()
    c = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    print(repr(t), c, ",".join(k for k, b in bits.items() if r & b) or "-", flush=True)
    time.sleep(0.004)
"""


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML from a child process (the bench thread holds the GIL
    while it queues launches, so an in-process sampler thread starves during a short timed region).  The child runs from
    before the warm-up; stop() keeps the samples whose time stamps fall inside [start(), stop()]."""

    def __init__(self, index: int):
        self.proc, self.t0, self.err = None, None, None
        try:
            import subprocess
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.t0 = time.time()

    def stop(self):
        t1 = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": self.err}
        time.sleep(0.02)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out, _ = self.proc.communicate()
        mx, inside, allc, reasons = None, [], [], set()
        for ln in out.splitlines():
            f = ln.split()
            if len(f) == 2 and f[0] == "max":
                mx = int(f[1])
            elif len(f) == 3:
                t, c = float(f[0]), int(f[1])
                allc.append((t, c))
                if self.t0 <= t <= t1:
                    inside.append(c)
                    if f[2] != "-":
                        reasons.update(f[2].split(","))
        if not inside and allc:  # region shorter than one sampling period: the sample nearest to its middle
            mid = 0.5 * (self.t0 + t1)
            inside = [min(allc, key=lambda tc: abs(tc[0] - mid))[1]]
        med = float(np.median(inside)) if inside else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(inside)}


def make_inputs(rank: int, n_img: int):
    from cic_b200 import synth
    first = rank * n_img
    img = synth.to_signed_range(synth.synth_images_u8(n_img, IMG_HW, IMG_HW, seed=synth.SEED_BASE + 1, first_index=first))
    mask = synth.synth_masks(n_img, IMG_HW, IMG_HW, seed=synth.SEED_BASE + 1, first_index=first)
    bpp = np.full((n_img, 1), TARGET_BPP, np.float32)
    return img, mask, bpp


def tiles_of(a, c):
    n = a.shape[0]
    t = IMG_HW // TILE
    return a.reshape(n, t, TILE, t, TILE, c).transpose(0, 1, 3, 2, 4, 5).reshape(-1, TILE, TILE, c)


def cpu_oracle_rate(weights, img, mask, bpp, n_tiles: int, reps: int = 1):
    """MPix/s of the CPU oracle (torch fp32, all host threads) on the first n_tiles tiles of the workload."""
    import torch
    from oracle import graphs, metrics
    torch.set_num_threads(os.cpu_count() or 1)
    ti, tm = tiles_of(img, 3)[:n_tiles], tiles_of(mask, 1)[:n_tiles]
    tb = np.repeat(bpp.reshape(-1), (IMG_HW // TILE) ** 2)[:n_tiles].reshape(-1, 1)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        outs = graphs.adaptive_forward(weights, ti, tm, tb)
        for k in range(min(n_tiles, 4)):
            metrics.compute_metrics(ti[k], outs[0][k])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_tiles * TILE * TILE / best / 1e6, best


def run_reference(args):
    """--impl reference: the reference's CPU path (its TensorFlow stack is not installable here, so the
    oracle port of the same graph) timed on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import importlib
    W = importlib.import_module("contextual-image-compression_b200.weights")
    import torch
    n_tiles = 8
    img, mask, bpp = make_inputs(0, n_tiles // 4)
    weights = W.synthetic_adaptive((TILE, TILE, 3), BASE_LATENT, seed=42)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_oracle_rate(weights, img, mask, bpp, n_tiles)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_rate(weights, img, mask, bpp, n_tiles)
    dt = (time.perf_counter() - t0) / args.steps
    val = n_tiles * TILE * TILE / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "adaptive GAN codec, 512x512 images as 256x256 tiles, bpp 1.0 (BASELINE configs[1])",
                       "sample": f"{n_tiles} tiles ({n_tiles // 4} images) per step", "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": val, "unit": "MPix/s", "cores": cores, "kind": "port",
                             "sample": f"{n_tiles} tiles of 256x256 per step, torch {torch.__version__} CPU fp32 oracle, {cores} threads"},
            "e2e": {"value": val, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CIC_PRECISION", "tc"), choices=["tc", "fp32"])
    ap.add_argument("--images", type=int, default=IMGS_PER_GPU, help="512x512 images per GPU per step")
    ap.add_argument("--e2e-chunks", default="auto", help="pipelined end-to-end leg: number of chunks of the batch, comma-separated chunk "
                    "sizes, or 'auto' (n/8, n/4, n/2, n/8: short first upload and last download)")
    ap.add_argument("--e2e-mode", default="phased", choices=["phased", "pipelined"],
                    help="end-to-end leg: predict_phased (encoder convs per upload chunk, Dense layers once per batch, decoders per "
                         "download chunk) or predict_pipelined (the whole graph per chunk)")
    ap.add_argument("--enc-chunks", default=None, help="phased: comma-separated upload / encode chunk sizes (default n/16, 3n/16, n/4, n/4, n/4)")
    ap.add_argument("--dec-chunks", default=None, help="phased: comma-separated decode / download chunk sizes (default: the encode schedule reversed)")
    ap.add_argument("--ssim", default="fast", choices=["fast", "exact"],
                    help="SSIM kernel of the evaluation: float32 window sums on centred data (HBM-bound, |dSSIM| < 1e-5 against "
                         "scikit-image in the tests) or the op-by-op scipy arithmetic (double accumulation, conversion-pipe-bound)")
    ap.add_argument("--cpu-tiles", type=int, default=8, help="tiles in the CPU-oracle sample (0 = skip)")
    ap.add_argument("--profile-csv", default=None, help="write the per-layer device times of the last timed step here")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import cic_b200 as cic
    import GAN_functions as gf

    rank, world = cic.dist.init()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local_rank if world > 1 else torch.cuda.current_device())
    torch.cuda.set_device(dev)
    numa_cpus = cic.dist.bind_to_gpu_numa_node(dev.index if dev.index is not None else 0) if world > 1 else 0
    cic.set_precision(args.precision)
    peaks = measured_peaks()
    # clock sampler child (started now so that it is up before the timed region); NVML indices are physical
    cuda_idx = dev.index if dev.index is not None else 0
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    if vis and all(v.strip().isdigit() for v in vis.split(",")) and cuda_idx < len(vis.split(",")):
        cuda_idx = int(vis.split(",")[cuda_idx])
    sampler = ClockSampler(cuda_idx)

    n_img = args.images
    n_tiles = n_img * (IMG_HW // TILE) ** 2
    img, mask, bpp = make_inputs(rank, n_img)
    weights = cic.weights.synthetic_adaptive((TILE, TILE, 3), BASE_LATENT, seed=42)
    models = gf.build_adaptive_compression_model((TILE, TILE, 3), BASE_LATENT, target_bpp=True)
    am = models["adaptive_model"]
    am.set_weights_dict(weights)

    # pinned host buffers (e2e leg) and device-resident copies (value leg)
    h_img, h_mask, h_bpp = (torch.from_numpy(a).pin_memory() for a in (img, mask, bpp))
    d_img, d_mask, d_bpp = (t.to(dev) for t in (h_img, h_mask, h_bpp))
    nfields = len(cic.dist.METRIC_FIELDS)
    px_per_step = n_img * IMG_HW * IMG_HW

    fast_ssim = args.ssim == "fast"

    def evaluate(d_in, outs):
        """bpp / PSNR / SSIM evaluation of the step + the one exchange step of the path."""
        m = cic.ops.metrics_f32(d_in, outs["blended"], signed_range=True, fast=fast_ssim)   # (n,4) psnr, ssim, mse, sse
        sums = cic.ops.metric_sums(m, outs["hq_ratio_sum"], IMG_HW * IMG_HW, 2 * BASE_LATENT, BASE_LATENT, TILE * TILE)
        return cic.dist.allreduce_metric_sums(sums)

    def step_device():
        outs = am.forward_device([d_img, d_mask, d_bpp], extras=False)
        return evaluate(d_img, am.last)

    if args.e2e_chunks == "auto":
        # graded chunks: a short first upload and last download, sizes doubling in between so every upload hides behind the
        # previous chunk's kernels (measured r01: 8,16,32,8 beats 8,48,8 and five- or six-chunk schedules at 64 images)
        if n_img >= 32:
            a, b = n_img // 8, n_img // 4
            e2e_chunks = [a, b, n_img - 2 * a - b, a]
        else:
            e2e_chunks = [n_img // 8, n_img - 2 * (n_img // 8), n_img // 8] if n_img >= 16 else min(n_img, 2)
    else:
        e2e_chunks = [int(v) for v in str(args.e2e_chunks).split(",")] if "," in str(args.e2e_chunks) else int(args.e2e_chunks)

    def evaluate_chunk(d_in, outs):
        """Per-chunk metric sums (no all-reduce): runs on the compute stream inside the pipelined predict."""
        k = d_in[0].shape[0]
        m = cic.ops.metrics_f32(d_in[0], outs["blended"], signed_range=True, fast=fast_ssim)
        return cic.ops.metric_sums(m, outs["hq_ratio_sum"], IMG_HW * IMG_HW, 2 * BASE_LATENT, BASE_LATENT, TILE * TILE)

    enc_chunks = [int(v) for v in args.enc_chunks.split(",")] if args.enc_chunks else None
    dec_chunks = [int(v) for v in args.dec_chunks.split(",")] if args.dec_chunks else None

    def step_e2e():
        # pinned host buffers -> (H2D | model + metrics | D2H of all 5 outputs) pipelined over chunks of the batch
        if args.e2e_mode == "phased":
            outs, parts = am.predict_phased([h_img, h_mask, h_bpp], enc_chunks=enc_chunks, dec_chunks=dec_chunks, on_chunk=evaluate_chunk)
        else:
            outs, parts = am.predict_pipelined([h_img, h_mask, h_bpp], n_chunks=e2e_chunks, on_chunk=evaluate_chunk)
        sums = cic.dist.allreduce_metric_sums(torch.stack(parts).sum(0))
        return outs, sums.cpu()

    plan = am.plan()
    launches_per_step = None

    # ---------------- value leg: inputs resident in HBM -------------------------------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    plan.set_profiling(True)
    cic.dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    layer_ms, layer_flops = {}, {}
    e0.record()
    for _ in range(args.steps):
        sums = step_device()
    e1.record()
    torch.cuda.synchronize()
    cic.dist.barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    ms_step = cic.dist.max_over_ranks(ms_total / args.steps, device=dev)
    prof = plan.profile()                       # per-layer device times of the last timed step
    plan.set_profiling(False)
    launches_per_step = plan.last_launch_count() + 3          # + metrics kernel, its finalise and the metric sums
    value = world * px_per_step / (ms_step * 1e-3) / 1e6

    KINDS = {0: "other", 1: "tc_gemm_kernel", 2: "tc_conv_kernel", 3: "conv1_tc_kernel", 4: "direct_conv_kernel", 5: "igemm_f32_kernel", 6: "conv_rows_tc_kernel", 7: "tc_gemm2_kernel", 8: "attn_fused_kernel"}
    by_kind = {}
    for name, ms, fl, by, kind in prof:
        k = by_kind.setdefault(KINDS.get(kind, "other"), {"ms": 0.0, "flops": 0.0, "launches": 0})
        k["ms"] += ms
        k["flops"] += fl
        k["launches"] += 1
    gemm_ms = sum(ms for name, ms, fl, by, kind in prof if fl > 0)
    gemm_flops = sum(fl for name, ms, fl, by, kind in prof if fl > 0)
    total_layer_ms = sum(ms for name, ms, fl, by, kind in prof)
    # dominant kernel = the kernel class with the largest share of the step
    dom = max((k for k in by_kind if k != "other"), key=lambda k: by_kind[k]["ms"]) if by_kind else "other"
    dom_ms, dom_flops = by_kind.get(dom, {"ms": 0.0})["ms"], by_kind.get(dom, {"flops": 0.0})["flops"]
    achieved_tflops = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    # executed tensor-core FLOPs of the same launches: the encoder chain (conv2-4, attention projections, Dense) issues three MMAs
    # per algorithmic MAC (hi*hi + lo*hi + hi*lo, split-bf16); reported next to the algorithmic figure, never instead of it
    dom_exec = sum(fl * (3.0 if (args.precision == "tc" and ("_enc/" in name)) else 1.0)
                   for name, ms, fl, by, kind in prof if KINDS.get(kind, "other") == dom)
    executed_tflops = dom_exec / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram bytes per step of each kernel class, from ncu --set full
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_step")
    roofline = {"bound": "tensor", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tflops / peak,
                "traffic": traffic,
                "kernel": f"{dom} ({by_kind.get(dom, {}).get('launches', 0)} launches per step; algorithmic FLOPs 2*M*N*K of its layers / "
                          f"their summed device time, CUDA events around every launch on the launch stream inside the timed steps)",
                "peak_source": f"{peaks['source']} sustained bf16 (MEASURED_PEAKS.json)",
                "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / ms_step if ms_step else None,
                "kernel_algorithmic_flops_per_step": dom_flops,
                "kernel_executed_tflops": executed_tflops, "kernel_executed_frac_of_burst_peak": executed_tflops / peaks["bf16_tflops"],
                "all_gemm_layers": {"tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0, "ms_per_step": gemm_ms,
                                    "algorithmic_flops_per_step": gemm_flops, "model_flops_per_step": n_tiles * FLOP_PER_TILE,
                                    "frac_of_peak": (gemm_flops / (gemm_ms * 1e-3) / 1e12 / peak) if gemm_ms > 0 else 0.0},
                "by_kernel": {k: {"ms": v["ms"], "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0,
                                  "launches": v["launches"]} for k, v in by_kind.items()},
                "note": "encoder layers execute 3 MMAs per algorithmic MAC (split-bf16); executed tensor FLOPs are higher than algorithmic"}
    if args.profile_csv and rank == 0:
        with open(args.profile_csv, "w") as f:
            f.write("layer,ms,flops,bytes,tflops,gbs,kernel\n")
            for name, ms, fl, by, kind in prof:
                f.write(f"{name},{ms:.4f},{fl:.4e},{by:.4e},{(fl / ms / 1e9) if ms > 0 else 0:.2f},{(by / ms / 1e6) if ms > 0 else 0:.1f},"
                        f"{KINDS.get(kind, 'other')}\n")

    # ---------------- e2e leg: pinned host buffers in, host buffers out ------------------------------
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    cic.dist.barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        outs, sums_host = step_e2e()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    ms_e2e = cic.dist.max_over_ranks(max(e0.elapsed_time(e1) / args.steps, wall), device=dev)
    e2e_value = world * px_per_step / (ms_e2e * 1e-3) / 1e6
    h2d = int(h_img.numel() * 4 + h_mask.numel() * 4 + h_bpp.numel() * 4)
    d2h = int(sum(o.nbytes for o in outs) + sums_host.numel() * 8)

    # informational: the same leg with the reference's uint8 image conventions applied on the device (uint8 image up, uint8 blended
    # image down: 1 instead of 4 bytes per sample over PCIe); the headline `e2e` above moves the float32 arrays of the reference API
    e2e_u8 = None
    if args.e2e_mode == "phased":
        from cic_b200 import synth
        h_img_u8 = torch.from_numpy(synth.synth_images_u8(n_img, IMG_HW, IMG_HW, seed=synth.SEED_BASE + 1, first_index=rank * n_img)).pin_memory()

        def step_u8():
            o, parts = am.predict_phased([h_img_u8, h_mask, h_bpp], enc_chunks=enc_chunks, dec_chunks=dec_chunks, on_chunk=evaluate_chunk, u8_io=True)
            return o, cic.dist.allreduce_metric_sums(torch.stack(parts).sum(0)).cpu()
        for _ in range(2):
            step_u8()
        torch.cuda.synchronize()
        cic.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            outs8, _s8 = step_u8()
        torch.cuda.synchronize()
        ms_u8 = cic.dist.max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3, device=dev)
        e2e_u8 = {"value": world * px_per_step / (ms_u8 * 1e-3) / 1e6, "unit": "MPix/s", "ms_per_step": ms_u8,
                  "h2d_bytes_per_step": int(h_img_u8.numel() + h_mask.numel() * 4 + h_bpp.numel() * 4),
                  "d2h_bytes_per_step": int(sum(o.nbytes for o in outs8) + 64), "api": "adaptive_model.predict_phased(u8_io=True)"}

    # ---------------- HBM-bound kernels of the path: algorithmic bytes / device time ----------------------
    hbm_peak = peaks["hbm_gbs"]
    hbm = {}
    for name, ms, fl, by, kind in prof:
        if name in ("roi_blend", "quantize") and ms > 0:
            hbm[name] = {"ms": ms, "algorithmic_bytes": by, "gbs": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / hbm_peak}
    am.forward_device([d_img, d_mask, d_bpp], extras=False)
    blended = am.last["blended"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    m_bytes = 24.0 * px_per_step                                  # two fp32 RGB images read once
    reps = 5
    for key, fast, note in (("metrics_psnr_ssim_f32_fast", True, "float32 window sums on centred data (DESIGN.md 4.5)"),
                            ("metrics_psnr_ssim_f32", False, "bound by the fp32<->fp64 conversion pipe, not HBM: scipy-exact 7x7 window "
                                                             "sums (double accumulation, float32 after each pass), DESIGN.md 4.5")):
        cic.ops.metrics_f32(d_img, blended, signed_range=True, fast=fast)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            mm = cic.ops.metrics_f32(d_img, blended, signed_range=True, fast=fast)
        ev[1].record()
        torch.cuda.synchronize()
        m_ms = ev[0].elapsed_time(ev[1]) / reps
        hbm[key] = {"ms": m_ms, "algorithmic_bytes": m_bytes, "gbs": m_bytes / m_ms / 1e6,
                    "frac_of_hbm_peak": m_bytes / m_ms / 1e6 / hbm_peak, "mean_ssim": float(mm[:, 1].mean().item()), "note": note,
                    "used_in_step": fast == fast_ssim}
    roofline["hbm_kernels"] = hbm
    roofline["hbm_peak_gbs"] = hbm_peak

    s = sums.cpu().numpy()[0]
    n_total = s[6]
    quality = {"psnr_db": s[0] / n_total, "ssim": s[1] / n_total, "mse": s[2] / n_total, "actual_bpp": s[3] / n_total,
               "hq_ratio": s[4] / n_total, "images": int(n_total)}

    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if args.cpu_tiles > 0:
        cores = os.cpu_count() or 1
        rate, secs = cpu_oracle_rate(weights, img, mask, bpp, args.cpu_tiles, reps=2)
        cpu = {"value": rate, "unit": "MPix/s", "cores": cores, "kind": "port",
               "sample": f"first {args.cpu_tiles} tiles of the workload, best of 2, {secs:.2f} s, torch-CPU fp32 oracle of the reference graph "
                         f"(the reference's TensorFlow stack is not installable here)"}
    line = {"metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (3-term split-bf16 encoder, bf16 decoder, fp32 accumulate)" if args.precision == "tc" else "f32",
            "data": "synthetic",
            "config": {"workload": f"adaptive GAN codec (build_adaptive_compression_model), {n_img} x {IMG_HW}x{IMG_HW} images per GPU "
                                   f"= {n_tiles} tiles of 256x256, target bpp {TARGET_BPP} (BASELINE configs[1])",
                       "precision": args.precision, "tiles_per_gpu": n_tiles, "base_latent_dim": BASE_LATENT,
                       "l2": f"inputs per step {h2d / 1e6:.0f} MB + activations >> 126 MB L2 (no flush needed)",
                       "step": "encode + quantise + decode + ROI blend + PSNR/SSIM/bpp evaluation + metric all-reduce",
                       "ssim": args.ssim, "cpu_affinity": f"{numa_cpus} CPUs local to the rank's GPU" if numa_cpus else "unchanged"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "MPix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "ms_per_step": ms_e2e, "api": f"adaptive_model.predict_{args.e2e_mode}"},
            "e2e_u8_io": e2e_u8, "gpu_launches": int(launches_per_step * args.steps), "roofline": roofline, "cpu_baseline": cpu, "quality": quality,
            "layers_ms_per_step": total_layer_ms}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
