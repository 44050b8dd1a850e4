"""Python faces of the stand-alone libcic operators (device tensors in, device tensors out)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, runtime
from .runtime import ptr, to_device_f32

_ACT = {None: _lib.ACT_NONE, "linear": _lib.ACT_NONE, "relu": _lib.ACT_RELU, "lrelu": _lib.ACT_LRELU02,
        "sigmoid": _lib.ACT_SIGMOID, "tanh": _lib.ACT_TANH}


def conv2d_same(x, kernel, bias=None, stride=1, activation=None, scale=None, shift=None) -> torch.Tensor:
    """Keras Conv2D(padding='same'); x (B,H,W,Cin), kernel (kh,kw,Cin,Cout)."""
    x, kernel = to_device_f32(x), to_device_f32(kernel)
    b, h, w, cin = x.shape
    kh, kw, cin2, cout = kernel.shape
    if cin != cin2:
        raise ValueError(f"kernel expects {cin2} input channels, input has {cin}")
    bias = to_device_f32(bias) if bias is not None else None
    scale = to_device_f32(scale) if scale is not None else None
    shift = to_device_f32(shift) if shift is not None else None
    ho, wo = -(-h // stride), -(-w // stride)
    y = torch.empty((b, ho, wo, cout), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib.cic_conv2d_nhwc_f32(ptr(x), ptr(kernel), ptr(bias), ptr(scale), ptr(shift), ptr(y), b, h, w, cin,
                                            cout, kh, kw, stride, _ACT[activation], runtime.stream_ptr()))
    return y


def conv2d_transpose_k4s2(x, kernel, bias=None, activation=None, scale=None, shift=None) -> torch.Tensor:
    """Keras Conv2DTranspose(kernel 4, stride 2, 'same'); kernel (4,4,Cout,Cin)."""
    x, kernel = to_device_f32(x), to_device_f32(kernel)
    b, h, w, cin = x.shape
    if tuple(kernel.shape[:2]) != (4, 4) or kernel.shape[3] != cin:
        raise ValueError(f"kernel must be (4,4,Cout,{cin}), got {tuple(kernel.shape)}")
    cout = kernel.shape[2]
    bias = to_device_f32(bias) if bias is not None else None
    scale = to_device_f32(scale) if scale is not None else None
    shift = to_device_f32(shift) if shift is not None else None
    y = torch.empty((b, 2 * h, 2 * w, cout), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib.cic_conv2d_transpose4x4s2_nhwc_f32(ptr(x), ptr(kernel), ptr(bias), ptr(scale), ptr(shift), ptr(y),
                                                           b, h, w, cin, cout, _ACT[activation], runtime.stream_ptr()))
    return y


def dense(x, kernel, bias=None, activation=None) -> torch.Tensor:
    x, kernel = to_device_f32(x), to_device_f32(kernel)
    b, k = x.shape
    if kernel.shape[0] != k:
        raise ValueError(f"kernel expects {kernel.shape[0]} inputs, got {k}")
    n = kernel.shape[1]
    bias = to_device_f32(bias) if bias is not None else None
    y = torch.empty((b, n), dtype=torch.float32, device=x.device)
    wsb = int(_lib.lib.cic_dense_workspace_bytes(b, k, n))
    ws = runtime.Workspace.get(wsb) if wsb else None
    _lib.check(_lib.lib.cic_dense_f32(ptr(x), ptr(kernel), ptr(bias), ptr(y), b, k, n, _ACT[activation], ptr(ws),
                                      ws.numel() if ws is not None else 0, runtime.stream_ptr()))
    return y


def conv2d_tc(x, kernel, bias=None, stride=1, activation=None, scale=None, shift=None, x2=None, transpose=False,
              split=False) -> torch.Tensor:
    """Tensor-core Conv2D('same') / Conv2DTranspose(k4, s2) on concat([x, x2]); see cic_conv2d_nhwc_tc."""
    x, kernel = to_device_f32(x), to_device_f32(kernel)
    b, h, w, cin = x.shape
    cin2 = 0
    if x2 is not None:
        x2 = to_device_f32(x2)
        if tuple(x2.shape[:3]) != (b, h, w):
            raise ValueError(f"x2 {tuple(x2.shape)} does not match x {tuple(x.shape)}")
        cin2 = x2.shape[3]
    kh, kw = int(kernel.shape[0]), int(kernel.shape[1])
    if transpose:
        if (kh, kw) != (4, 4) or kernel.shape[3] != cin + cin2:
            raise ValueError(f"kernel must be (4,4,Cout,{cin + cin2}), got {tuple(kernel.shape)}")
        cout, stride = int(kernel.shape[2]), 2
        y = torch.empty((b, 2 * h, 2 * w, cout), dtype=torch.float32, device=x.device)
    else:
        if kernel.shape[2] != cin + cin2:
            raise ValueError(f"kernel expects {kernel.shape[2]} input channels, inputs have {cin + cin2}")
        cout = int(kernel.shape[3])
        y = torch.empty((b, -(-h // stride), -(-w // stride), cout), dtype=torch.float32, device=x.device)
    bias = to_device_f32(bias) if bias is not None else None
    scale = to_device_f32(scale) if scale is not None else None
    shift = to_device_f32(shift) if shift is not None else None
    wsb = int(_lib.lib.cic_conv2d_tc_workspace_bytes(b, h, w, cin, cin2, cout, kh, kw, stride, int(transpose)))
    ws = runtime.Workspace.get(wsb)
    _lib.check(_lib.lib.cic_conv2d_nhwc_tc(ptr(x), ptr(x2), ptr(kernel), ptr(bias), ptr(scale), ptr(shift), ptr(y), b, h, w,
                                           cin, cin2, cout, kh, kw, stride, int(transpose), _ACT[activation], int(split),
                                           ptr(ws), ws.numel(), runtime.stream_ptr()))
    return y


def dense_tc(x, kernel, bias=None, activation=None, split=False) -> torch.Tensor:
    x, kernel = to_device_f32(x), to_device_f32(kernel)
    b, k = x.shape
    n = kernel.shape[1]
    bias = to_device_f32(bias) if bias is not None else None
    y = torch.empty((b, n), dtype=torch.float32, device=x.device)
    wsb = int(_lib.lib.cic_dense_tc_workspace_bytes(b, k, n))
    ws = runtime.Workspace.get(wsb)
    _lib.check(_lib.lib.cic_dense_tc(ptr(x), ptr(kernel), ptr(bias), ptr(y), b, k, n, _ACT[activation], int(split), ptr(ws),
                                     ws.numel(), runtime.stream_ptr()))
    return y


def self_attention(x, wq, bq, wk, bk, wv, bv, gamma: float) -> torch.Tensor:
    """SelfAttention.call (GAN_functions.py:344-369); x (B,H,W,C); wq/wk (1,1,C,C/8) or (C,C/8); wv (.,C,C)."""
    x = to_device_f32(x)
    b, h, w, c = x.shape
    wq, wk, wv = (to_device_f32(t).reshape(c, -1) for t in (wq, wk, wv))
    bq, bk, bv = (to_device_f32(t) if t is not None else None for t in (bq, bk, bv))
    y = torch.empty_like(x)
    wsb = int(_lib.lib.cic_attention_workspace_bytes(b, h * w, c))
    ws = runtime.Workspace.get(wsb)
    _lib.check(_lib.lib.cic_self_attention_f32(ptr(x), ptr(wq), ptr(bq), ptr(wk), ptr(bk), ptr(wv), ptr(bv), float(gamma),
                                               ptr(y), b, h * w, c, ptr(ws), ws.numel(), runtime.stream_ptr()))
    return y


def quantize_latent(latent, saliency, quant_strength, want=("deq",)):
    """AdaptiveQuantizationLayer.call (GAN_functions.py:435-446).  Returns a dict with the requested
    outputs among 'deq', 'symbols', 'pre', 'scale'."""
    latent = to_device_f32(latent)
    b, L = latent.shape
    sal = to_device_f32(saliency).reshape(-1)
    qs = to_device_f32(quant_strength).reshape(-1)
    if sal.numel() != b or qs.numel() != b:
        raise ValueError(f"saliency and quant_strength need one value per row ({b}), got {sal.numel()} and {qs.numel()}")
    dev = latent.device
    out = {}
    if "deq" in want:
        out["deq"] = torch.empty_like(latent)
    if "symbols" in want:
        out["symbols"] = torch.empty((b, L), dtype=torch.int32, device=dev)
    if "pre" in want:
        out["pre"] = torch.empty_like(latent)
    if "scale" in want:
        out["scale"] = torch.empty((b,), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib.cic_quantize_latent(ptr(latent), ptr(sal), ptr(qs), ptr(out.get("deq")), ptr(out.get("symbols")),
                                            ptr(out.get("pre")), ptr(out.get("scale")), b, L, runtime.stream_ptr()))
    return out


def rate_scalars(target_bpp):
    """t, hq_lq_threshold, quant_strength of GAN_functions.py:631-649."""
    bpp = to_device_f32(target_bpp).reshape(-1)
    n = bpp.numel()
    t, thr, qs = (torch.empty_like(bpp) for _ in range(3))
    _lib.check(_lib.lib.cic_rate_scalars(ptr(bpp), ptr(t), ptr(thr), ptr(qs), n, runtime.stream_ptr()))
    return t, thr, qs


def roi_mask_blend(hq, lq, mask, target_bpp, want_dt=True, want_sum=True):
    """dt = sigmoid((mask^0.7 - thr)*20); out = hq*dt + lq*(1-dt) (GAN_functions.py:651-684).
    hq/lq may be None to get only the bit-allocation map."""
    mask = to_device_f32(mask)
    b = mask.shape[0]
    hw = int(np.prod(mask.shape[1:3]))
    bpp = to_device_f32(target_bpp).reshape(-1)
    if bpp.numel() != b:
        raise ValueError("one target bpp per image")
    dev = mask.device
    out = None
    c = 3
    if hq is not None:
        hq, lq = to_device_f32(hq), to_device_f32(lq)
        if hq.shape != lq.shape or hq.shape[0] != b or int(np.prod(hq.shape[1:3])) != hw:
            raise ValueError(f"hq {tuple(hq.shape)}, lq {tuple(lq.shape)} and mask {tuple(mask.shape)} do not agree")
        c = hq.shape[-1]
        out = torch.empty_like(hq)
    dt = torch.empty(mask.shape, dtype=torch.float32, device=dev) if want_dt else None
    s = torch.empty((b,), dtype=torch.float64, device=dev) if want_sum else None
    _lib.check(_lib.lib.cic_roi_mask_blend(ptr(hq), ptr(lq), ptr(mask), ptr(bpp), ptr(out), ptr(dt), ptr(s), b, hw, c,
                                           runtime.stream_ptr()))
    return out, dt, s


def hq_ratio_sweep(mask, bpp_levels) -> torch.Tensor:
    """(B, n_levels) float64 hq_ratio = mean(dt) for every image x target bpp, one pass over the masks."""
    mask = to_device_f32(mask)
    b = mask.shape[0]
    hw = int(np.prod(mask.shape[1:3]))
    levels = to_device_f32(np.asarray(bpp_levels, dtype=np.float32) if not isinstance(bpp_levels, torch.Tensor) else bpp_levels).reshape(-1)
    out = torch.empty((b, levels.numel()), dtype=torch.float64, device=mask.device)
    _lib.check(_lib.lib.cic_hq_ratio_sweep(ptr(mask), ptr(levels), levels.numel(), ptr(out), b, hw, runtime.stream_ptr()))
    return out


def symbol_entropy_bits(symbols: torch.Tensor) -> torch.Tensor:
    if symbols.dtype != torch.int32 or not symbols.is_cuda:
        symbols = torch.as_tensor(np.asarray(symbols.cpu() if isinstance(symbols, torch.Tensor) else symbols)).to(torch.int32).to(runtime.require_cuda())
    symbols = symbols.contiguous()
    b, L = symbols.shape
    out = torch.empty((b,), dtype=torch.float64, device=symbols.device)
    _lib.check(_lib.lib.cic_symbol_entropy_bits(ptr(symbols), ptr(out), b, L, runtime.stream_ptr()))
    return out


def saliency_mask_smooth(saliency_map) -> torch.Tensor:
    """create_saliency_mask(saliency_map, smooth=True) (GAN_functions.py:199-203) on the device: bilateral(9, 75, 75) -> Gaussian
    31x31 -> / max, OpenCV semantics.  saliency_map (H,W) or (B,H,W) -> float32 mask of the same shape on the device."""
    x = to_device_f32(saliency_map)
    single = x.dim() == 2
    if single:
        x = x.unsqueeze(0)
    if x.dim() != 3:
        raise ValueError(f"saliency map must be (H,W) or (B,H,W), got {tuple(x.shape)}")
    b, h, w = x.shape
    y = torch.empty_like(x)
    ws = torch.empty(int(_lib.lib.cic_saliency_mask_workspace_bytes(b, h, w)), dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib.cic_saliency_mask_smooth(ptr(x), ptr(y), b, h, w, ptr(ws), ws.numel(), runtime.stream_ptr()))
    return y[0] if single else y


def saliency_enhance(saliency_map) -> torch.Tensor:
    """enhance_saliency_map(saliency_map) (GAN_functions.py:123-157) on the device: bilateral(9, 75, 75) -> Gaussian 3 / 9 / 15 mixed
    0.5 / 0.3 / 0.2 -> ^0.8 -> clip.  (H,W) or (B,H,W) -> float32 of the same shape."""
    x = to_device_f32(saliency_map)
    single = x.dim() == 2
    if single:
        x = x.unsqueeze(0)
    if x.dim() != 3:
        raise ValueError(f"saliency map must be (H,W) or (B,H,W), got {tuple(x.shape)}")
    b, h, w = x.shape
    y = torch.empty_like(x)
    ws = torch.empty(int(_lib.lib.cic_saliency_enhance_workspace_bytes(b, h, w)), dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib.cic_saliency_enhance(ptr(x), ptr(y), b, h, w, ptr(ws), ws.numel(), runtime.stream_ptr()))
    return y[0] if single else y


def saliency_mask_binary(saliency_map, threshold=None, return_threshold: bool = False):
    """create_saliency_mask(saliency_map, threshold, smooth=False) (GAN_functions.py:172-197, :204-206) on the device: the map
    compared with `threshold`, or with the reference's adaptive threshold (OpenCV's Otsu on the uint8 map vs the 70 % share of a
    50-bin histogram, clamped to [0.05, 0.5]) when none is given.  (H,W) or (B,H,W) -> float32 0 / 1 mask of the same shape
    [, float64 thresholds (B,)]."""
    x = to_device_f32(saliency_map)
    single = x.dim() == 2
    if single:
        x = x.unsqueeze(0)
    if x.dim() != 3:
        raise ValueError(f"saliency map must be (H,W) or (B,H,W), got {tuple(x.shape)}")
    b, h, w = x.shape
    y = torch.empty_like(x)
    adaptive = threshold is None
    thr = torch.empty((b,), dtype=torch.float64, device=x.device) if adaptive else None
    ws = torch.empty(int(_lib.lib.cic_saliency_mask_binary_workspace_bytes(b)) if adaptive else 0, dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib.cic_saliency_mask_binary(ptr(x), ptr(y), 0.0 if adaptive else float(threshold), int(adaptive), ptr(thr) if adaptive else None,
                                                 b, h, w, ptr(ws) if adaptive else None, ws.numel(), runtime.stream_ptr()))
    if not adaptive:
        thr = torch.full((b,), float(threshold), dtype=torch.float64, device=x.device)
    y = y[0] if single else y
    return (y, thr) if return_threshold else y


SALIENCY_METHODS = {"spectral_residual": 0, "fine_grained": 1, "combined": 2}      # enum cic_saliency_method


def saliency_map(image, method: str = "spectral_residual") -> torch.Tensor:
    """compute_saliency_map(image, method) (GAN_functions.py:52-121) on the device: the spectral-residual and fine-grained
    detectors of cv2.saliency (opencv-contrib) and their 0.6 / 0.4 mix, divided by the maximum.  image: (H,W,3) or (B,H,W,3), RGB;
    float32 with max <= 1 is taken as [-1, 1] and mapped with ((x + 1) * 127.5).astype(uint8), anything else is cast to uint8
    (:63-67).  -> float32 map (H,W) or (B,H,W) on the device."""
    if method not in SALIENCY_METHODS:
        raise ValueError(f"Unsupported saliency method: {method}")                      # GAN_functions.py:110
    x = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image))
    single = x.dim() == 3
    if single:
        x = x[None]
    if x.dim() != 4 or x.shape[3] != 3:
        raise ValueError(f"saliency_map: expected (H,W,3) or (B,H,W,3), got {tuple(x.shape)}")
    x = x.to(runtime.require_cuda())
    if x.dtype == torch.float32 and x.numel() and float(x.max()) <= 1.0:
        x = f32_signed_to_u8(x.contiguous())
    elif x.dtype != torch.uint8:
        x = x.to(torch.int32).to(torch.uint8)           # numpy's astype(uint8): truncate toward zero, wrap modulo 256
    x = x.contiguous()
    b, h, w, _ = x.shape
    y = torch.empty((b, h, w), dtype=torch.float32, device=x.device)
    ws = torch.empty(int(_lib.lib.cic_saliency_map_workspace_bytes(b, h, w)), dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib.cic_saliency_map_u8(ptr(x), ptr(y), b, h, w, 1, SALIENCY_METHODS[method], ptr(ws), ws.numel(),
                                            runtime.stream_ptr()))
    return y[0] if single else y


def saliency_mask_from_image(image, method: str = "combined") -> torch.Tensor:
    """The reference's whole mask front end for one image or a batch, on the device with nothing crossing PCIe in between:
    create_saliency_mask(compute_saliency_map(img, method), smooth=True) (GAN_test.py:279-280)."""
    return saliency_mask_smooth(saliency_map(image, method))


def rans_encode(symbols: torch.Tensor):
    """Entropy-code integer latent symbols (rows, L) int32 on the device -> (uint8 stream tensor, trimmed to its length).  One
    host synchronisation (the length).  Symbols beyond +-1023 are clamped (|symbol| <~ 100 in the codec: scale <= e^2.652)."""
    if symbols.dtype != torch.int32 or not symbols.is_cuda:
        symbols = torch.as_tensor(np.asarray(symbols.cpu() if isinstance(symbols, torch.Tensor) else symbols)).to(torch.int32).to(runtime.require_cuda())
    symbols = symbols.contiguous()
    rows, L = symbols.shape
    dev = symbols.device
    cap = int(_lib.lib.cic_rans_max_bytes(rows, L))
    stream = torch.empty(cap, dtype=torch.uint8, device=dev)
    nbytes = torch.zeros((), dtype=torch.int64, device=dev)
    ws = torch.empty(int(_lib.lib.cic_rans_workspace_bytes(rows, L)), dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib.cic_rans_encode(ptr(symbols), rows, L, ptr(stream), cap, ptr(nbytes), ptr(ws), ws.numel(), runtime.stream_ptr()))
    return stream[: int(nbytes.item())]


def rans_decode(stream: torch.Tensor, rows: int, latent_dim: int) -> torch.Tensor:
    """Inverse of rans_encode -> (rows, latent_dim) int32 on the device."""
    if not isinstance(stream, torch.Tensor):
        stream = torch.from_numpy(np.ascontiguousarray(np.frombuffer(bytes(stream), dtype=np.uint8)))
    stream = stream.to(runtime.require_cuda()).contiguous()
    if stream.data_ptr() % 4:
        stream = stream.clone()
    if stream.numel() < 32 + 4096 + 4 * (rows + 1):
        raise ValueError(f"stream of {stream.numel()} bytes is shorter than the header of a {rows}-row stream")
    header = stream[:32].cpu().numpy().view("<u4")          # validated on the host: the kernel refuses a mismatching stream too
    if header[0] != 0x52434943 or header[1] != 1 or (int(header[2]), int(header[3])) != (rows, latent_dim):
        raise ValueError(f"not a CICR v1 stream of shape ({rows}, {latent_dim}): header says magic {int(header[0]):#x}, version {int(header[1])}, "
                         f"shape ({int(header[2])}, {int(header[3])})")
    out = torch.empty((rows, latent_dim), dtype=torch.int32, device=stream.device)
    _lib.check(_lib.lib.cic_rans_decode(ptr(stream), stream.numel(), ptr(out), rows, latent_dim, runtime.stream_ptr()))
    return out


def jpeg_encode_device(images_u8: torch.Tensor, quality: int = 95, rgb: bool = False, capacity: int = None):
    """Baseline JPEG files of a uint8 batch (B,H,W,3) on the device -> (out (B, capacity) uint8, sizes (B,) int32), both on the device;
    file b is out[b, :sizes[b]].  No host synchronisation.  A size above `capacity` means that file was truncated (cic.h)."""
    x = images_u8
    if x.dtype != torch.uint8 or not x.is_cuda or x.dim() != 4 or x.shape[3] != 3:
        raise ValueError(f"jpeg_encode_device: expected a (B,H,W,3) uint8 CUDA tensor, got {tuple(x.shape)} {x.dtype} on {x.device}")
    x = x.contiguous()
    b, h, w, _ = x.shape
    if capacity is None:
        capacity = int(_lib.lib.cic_jpeg_max_bytes(h, w))
    out = torch.empty((b, capacity), dtype=torch.uint8, device=x.device)
    sizes = torch.zeros((b,), dtype=torch.int32, device=x.device)
    ws = torch.empty(int(_lib.lib.cic_jpeg_workspace_bytes(b, h, w)), dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib.cic_jpeg_encode_u8(ptr(x), b, h, w, int(bool(rgb)), int(quality), ptr(out), capacity, ptr(sizes), ptr(ws), ws.numel(),
                                           runtime.stream_ptr()))
    return out, sizes


def jpeg_encode(images_u8, quality: int = 95, rgb: bool = False):
    """The bytes cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality]) returns (= the file cv2.imwrite writes:
    test_autoencoder.py:93, GAN_functions.py:50), made on the GPU.  images_u8: (H,W,3) or (B,H,W,3) uint8, BGR like cv2's input
    (rgb=True: RGB input, i.e. save_image's cvtColor folded in); device tensor or host array.  -> bytes or list of bytes.
    First tries a 4 bytes / pixel buffer (q95 photographs need ~1), and repeats with the strict worst case if a file did not fit."""
    x = images_u8 if isinstance(images_u8, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images_u8))
    single = x.dim() == 3
    if single:
        x = x[None]
    x = x.to(runtime.require_cuda())
    b, h, w, _ = x.shape
    worst = int(_lib.lib.cic_jpeg_max_bytes(h, w))
    cap = min(worst, 1024 + 4 * ((h + 15) // 16) * ((w + 15) // 16) * 256)
    out, sizes = jpeg_encode_device(x, quality, rgb, cap)
    n = sizes.cpu().numpy()
    if int(n.max()) > cap:
        out, sizes = jpeg_encode_device(x, quality, rgb, worst)
        n = sizes.cpu().numpy()
    top = int(n.max())
    host = out[:, :top].cpu().numpy()
    files = [host[i, : int(n[i])].tobytes() for i in range(b)]
    return files[0] if single else files


def f32_signed_to_u8(x) -> torch.Tensor:
    """((x + 1) * 127.5).astype(uint8) on the device (GAN_functions.py:44, save_image's first line)."""
    x = to_device_f32(x)
    y = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib.cic_f32_signed_to_u8(ptr(x), ptr(y), x.numel(), runtime.stream_ptr()))
    return y


def f32_to_u8_trunc(x, mul: float = 255.0) -> torch.Tensor:
    x = to_device_f32(x)
    y = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib.cic_f32_to_u8_trunc(ptr(x), ptr(y), x.numel(), float(mul), runtime.stream_ptr()))
    return y


def metrics_f32(a, b, signed_range: bool, data_range: float = 1.0, fast: bool = False) -> torch.Tensor:
    """(B,4) float64 [psnr, ssim, mse, sse]; signed_range maps [-1,1] -> [0,1] first (compute_metrics).
    fast=True: SSIM window sums in float32 on centred data (HBM-bound; ssim within ~1e-6 of scikit-image instead of the
    double-accumulating scipy arithmetic reproduced op by op); psnr / mse / sse are the same either way."""
    a, b = to_device_f32(a), to_device_f32(b)
    if a.dim() == 3:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    n, h, w, c = a.shape
    out = torch.empty((n, 4), dtype=torch.float64, device=a.device)
    pre_add, pre_mul = (1.0, 0.5) if signed_range else (0.0, 1.0)
    fn = _lib.lib.cic_metrics_psnr_ssim_f32_fast if fast else _lib.lib.cic_metrics_psnr_ssim_f32
    _lib.check(fn(ptr(a), ptr(b), ptr(out), n, h, w, c, pre_add, pre_mul, float(data_range), runtime.stream_ptr()))
    return out


def ms_ssim_f32(a, b, signed_range: bool, data_range: float = 1.0) -> torch.Tensor:
    """(B,) float64 MS-SSIM per image (five scales, 11x11 Gaussian window; see cic_msssim_f32 - an extra with no reference call
    site, BASELINE configs[4]).  Images (B,H,W,C) with H, W >= 176; signed_range maps [-1,1] -> [0,1] first."""
    a, b = to_device_f32(a), to_device_f32(b)
    if a.dim() == 3:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    n, h, w, c = a.shape
    if min(h, w) < 176:
        raise ValueError(f"MS-SSIM with five scales needs H, W >= 176, got {h}x{w}")
    out = torch.empty((n,), dtype=torch.float64, device=a.device)
    pre_add, pre_mul = (1.0, 0.5) if signed_range else (0.0, 1.0)
    ws = torch.empty(int(_lib.lib.cic_msssim_workspace_bytes(n, h, w, c)), dtype=torch.uint8, device=a.device)
    _lib.check(_lib.lib.cic_msssim_f32(ptr(a), ptr(b), ptr(out), n, h, w, c, pre_add, pre_mul, float(data_range), ptr(ws), ws.numel(),
                                       runtime.stream_ptr()))
    return out


def metric_sums(metrics: torch.Tensor, dt_sum: torch.Tensor, img_px: int, latent_hq: int, latent_lq: int, tile_px: int) -> torch.Tensor:
    """(1, 8) float64 row [sum psnr, sum ssim, sum mse, sum actual_bpp, sum hq_ratio, 0, n, 0] of one evaluated batch (dist.METRIC_FIELDS)
    from metrics_f32's (n,4) output and the per-image dt sums, in one launch (bpp accounting of GAN_test.py:310-325)."""
    n = metrics.shape[0]
    if metrics.dtype != torch.float64 or dt_sum.dtype != torch.float64 or dt_sum.numel() != n:
        raise ValueError("metric_sums takes the float64 outputs of metrics_f32 (n,4) and the per-image dt sums (n,)")
    out = torch.empty((1, 8), dtype=torch.float64, device=metrics.device)
    _lib.check(_lib.lib.cic_metric_sums(ptr(metrics.contiguous()), ptr(dt_sum.contiguous()), n, int(img_px), int(latent_hq), int(latent_lq),
                                        int(tile_px), ptr(out), runtime.stream_ptr()))
    return out


def metrics_gray_u8(a, b) -> torch.Tensor:
    """(B,4) float64 [psnr, ssim(gray), true mse, wrapped uint8 mse] for uint8 BGR images."""
    dev = runtime.require_cuda()
    ta = torch.as_tensor(a).to(dev).contiguous()
    tb = torch.as_tensor(b).to(dev).contiguous()
    if ta.dtype != torch.uint8 or tb.dtype != torch.uint8:
        raise ValueError("uint8 images expected")
    if ta.dim() == 3:
        ta, tb = ta.unsqueeze(0), tb.unsqueeze(0)
    n, h, w, c = ta.shape
    if c != 3 or ta.shape != tb.shape:
        raise ValueError(f"expected matching (B,H,W,3) images, got {tuple(ta.shape)} and {tuple(tb.shape)}")
    out = torch.empty((n, 4), dtype=torch.float64, device=dev)
    _lib.check(_lib.lib.cic_metrics_psnr_ssim_gray_u8(ptr(ta), ptr(tb), ptr(out), n, h, w, runtime.stream_ptr()))
    return out
