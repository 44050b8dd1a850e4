"""Device plumbing: PyTorch is only the container for device memory and streams."""
from __future__ import annotations

import os
from typing import Dict, Sequence

import numpy as np
import torch

from . import _lib

_PRECISION = {"fp32": _lib.PREC_FP32, "tc": _lib.PREC_TC}
_default_precision = os.environ.get("CIC_PRECISION", "tc").lower()


def set_precision(name: str) -> None:
    """'fp32' = fp32 CUDA-core arithmetic; 'tc' = tcgen05 tensor cores (split-bf16 encoder, bf16 decoder)."""
    global _default_precision
    if name not in _PRECISION:
        raise ValueError(f"precision must be one of {sorted(_PRECISION)}")
    _default_precision = name


def get_precision() -> str:
    return _default_precision


def precision_code(name: str | None = None) -> int:
    return _PRECISION[name or _default_precision]


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("cic_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_device_f32(x, device=None) -> torch.Tensor:
    """numpy / torch (any device) -> contiguous float32 CUDA tensor (no copy when already one)."""
    device = device or require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


class Workspace:
    """Grow-only per-device scratch buffer handed to the library as `d_workspace`."""

    _bufs: Dict[int, torch.Tensor] = {}

    @classmethod
    def get(cls, nbytes: int) -> torch.Tensor:
        dev = torch.cuda.current_device()
        buf = cls._bufs.get(dev)
        if buf is None or buf.numel() < nbytes:
            cls._bufs[dev] = buf = None  # release before growing
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=torch.device("cuda", dev))
            cls._bufs[dev] = buf
        return buf

    @classmethod
    def clear(cls) -> None:
        cls._bufs.clear()


_use_graphs = True
_pipe_timeline = False


def set_cuda_graphs(on: bool) -> None:
    """predict_pipelined / predict_phased replay CUDA graphs from their third call with one chunking (default on)."""
    global _use_graphs
    _use_graphs = bool(on)


def use_cuda_graphs() -> bool:
    return _use_graphs


def set_pipe_timeline(on: bool) -> None:
    """Print the per-chunk event times of the copy-in / compute / copy-out streams of the pipelined predicts (debugging)."""
    global _pipe_timeline
    _pipe_timeline = bool(on)


def pipe_timeline() -> bool:
    return _pipe_timeline


def as_list(x) -> list:
    if isinstance(x, (list, tuple)):
        return list(x)
    return [x]


def shapes_str(xs: Sequence) -> str:
    return ", ".join(str(tuple(getattr(x, "shape", ()))) for x in xs)
