"""Batch sharding across one-process-per-GPU ranks + the single exchange step of the path.

Images (autoencoder) and 256x256 tiles (GAN codec) are independent, so each rank codes a contiguous
slice of the global batch with replicated weights and no data-path collective.  The only exchange is an
all-reduce of the per-level metric sums (< 1 KB, latency-bound) and an optional
all-gather of per-image metrics.  Works with backend 'nccl' on GPUs and 'gloo' on CPU (tests).
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: str | None = None) -> Tuple[int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world


def bind_to_gpu_numa_node(cuda_index: int) -> int:
    """Pin this process to the CPUs NVML reports as local to its GPU (one process per GPU): pinned host buffers allocated
    afterwards are first-touched on that NUMA node, so every rank's PCIe copies stay on its own socket.  Returns the number of CPUs
    in the new affinity mask (0 = left unchanged: NVML or sched_setaffinity not available)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = cuda_index
        if vis and all(v.strip().isdigit() for v in vis.split(",")) and cuda_index < len(vis.split(",")):
            idx = int(vis.split(",")[cuda_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & allowed
        if not cpus or cpus == allowed:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001  (affinity is an optimisation, never a requirement)
        return 0


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global index range owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


METRIC_FIELDS = ("psnr_sum", "ssim_sum", "mse_sum", "actual_bpp_sum", "hq_ratio_sum", "entropy_bits_sum", "n",
                 "mismatch_count")


def metric_sums_row(psnr, ssim, mse, actual_bpp, hq_ratio, n_images: int) -> torch.Tensor:
    """(1, len(METRIC_FIELDS)) float64 row of per-rank sums built on the device with kernels only (no host scalar is
    copied into a device element), so it can sit inside a CUDA-graph capture."""
    dev = psnr.device
    f64 = dict(dtype=torch.float64, device=dev)
    zero = torch.zeros((), **f64)
    vals = [psnr.sum().to(torch.float64), ssim.sum().to(torch.float64), mse.sum().to(torch.float64),
            actual_bpp.sum().to(torch.float64), hq_ratio.sum().to(torch.float64), zero,
            torch.full((), float(n_images), **f64), zero]
    return torch.stack(vals).reshape(1, len(METRIC_FIELDS))


def allreduce_metric_sums(local: torch.Tensor) -> torch.Tensor:
    """Sum a (levels, len(METRIC_FIELDS)) float64 tensor over ranks (in place); identity for world 1."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
    return local


def allgather_per_image(local: torch.Tensor, counts) -> torch.Tensor:
    """Gather per-image rows (n_local, k) from every rank into global image order."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    mx = max(counts)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
