"""Keras-like model objects over libcic plans.

The reference's callers use two protocols:
  * `model.predict(x_or_list, verbose=0)` -> numpy array / list of numpy arrays
    (`test_autoencoder.py:85`, `GAN_test.py:292,567`);
  * `model(list, training=False)` -> sequence whose elements support `[0]` and `.numpy()`
    (`GAN_functions.py:867-871`) - CPU torch tensors satisfy that.
`forward_device` is the zero-copy entry point (CUDA tensors in, CUDA tensors out) used by the
benchmark's device-resident leg and by the multi-GPU runner.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import os
import sys
import numpy as np
import torch

from . import _lib, runtime, weights as W
from .runtime import ptr, to_device_f32


def _make_tensor_array(named: Dict[str, np.ndarray]):
    keep = []
    arr = (_lib.cic_tensor * max(len(named), 1))()
    for i, (name, a) in enumerate(named.items()):
        a32 = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a32)
        bname = name.encode()
        keep.append(bname)
        arr[i].name = bname
        arr[i].h_data = a32.ctypes.data_as(C.POINTER(C.c_float))
        arr[i].ndim = a32.ndim
        for d in range(a32.ndim):
            arr[i].shape[d] = a32.shape[d]
    return arr, keep


def phased_default_chunks(n: int) -> List[int]:
    """Upload / encode chunk sizes of predict_phased for a batch of n images (the decode / download schedule is the reverse).
    Measured r01 at 64 images: 4,12,16,16,16 / 16,16,16,12,4 beats coarser and finer schedules - a short first upload and a short
    last download are what stays exposed, equal chunks in between keep both copy engines busy."""
    if n >= 16:
        a = n // 16
        return [a, 3 * a, 4 * a, 4 * a, n - 12 * a]
    if n >= 4:
        return [n // 4, n // 4, n - 2 * (n // 4)]
    return [n]


class Plan:
    """RAII wrapper of a cic_plan*."""

    def __init__(self, kind: int, named_weights: Dict[str, np.ndarray], img_shape=(0, 0, 3), latent_dim=0,
                 add_attention=False, precision: Optional[str] = None):
        runtime.require_cuda()
        opts = _lib.cic_plan_opts()
        opts.precision = runtime.precision_code(precision)
        opts.img_h, opts.img_w, opts.img_c = int(img_shape[0]), int(img_shape[1]), int(img_shape[2])
        opts.latent_dim = int(latent_dim)
        opts.add_attention = int(bool(add_attention))
        arr, keep = _make_tensor_array(named_weights)
        self.handle = _lib.lib.cic_plan_create(kind, arr, len(named_weights), C.byref(opts))
        del keep
        if not self.handle:
            raise _lib.CicError(_lib.ERR_INVALID, _lib.last_error())
        self.kind = kind
        self.precision = precision or runtime.get_precision()

    def workspace_bytes(self, batch: int, h: int = 0, w: int = 0) -> int:
        return int(_lib.lib.cic_plan_workspace_bytes(self.handle, batch, h, w))

    def workspace(self, batch: int, h: int = 0, w: int = 0) -> torch.Tensor:
        """The shared grow-only scratch buffer (eager calls).  Captured CUDA graphs never use it: a graph bakes the raw
        pointer in, and the shared buffer is re-allocated when a later call needs more - they own private scratch instead."""
        return runtime.Workspace.get(self.workspace_bytes(batch, h, w))

    def last_launch_count(self) -> int:
        return int(_lib.lib.cic_plan_last_launch_count(self.handle))

    def set_profiling(self, on: bool) -> None:
        _lib.check(_lib.lib.cic_plan_set_profiling(self.handle, int(on)))

    def profile(self):
        """[(layer, ms, flops, bytes, kernel_kind)] of the last forward call (needs set_profiling(True) before it)."""
        n = int(_lib.lib.cic_plan_get_profile(self.handle, None, 0))
        buf = C.create_string_buffer(n + 16)
        _lib.lib.cic_plan_get_profile(self.handle, buf, n + 16)
        rows = []
        for line in buf.value.decode().splitlines():
            name, ms, fl, by, kind = line.rsplit(",", 4)
            rows.append((name, float(ms), float(fl), float(by), int(kind)))
        return rows

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                _lib.lib.cic_plan_destroy(h)
            except Exception:
                pass
            self.handle = None


class Model:
    """Common predict / __call__ protocol; subclasses implement forward_device."""

    name = "model"

    def __init__(self):
        self._plan: Optional[Plan] = None
        self._plan_precision: Optional[str] = None
        self._plan_gen = 0          # bumped whenever the plan is rebuilt: keys every cache that holds plan pointers
        self._ws_override: Optional[torch.Tensor] = None   # private scratch of the graph set being captured / replayed
        self._weights: Optional[Dict[str, np.ndarray]] = None
        self._default_weights = None   # () -> Keras-default weights, drawn on first use (a checkpoint usually replaces them first)

    # -- weights ------------------------------------------------------------------------------
    @property
    def weights(self) -> Dict[str, np.ndarray]:
        if self._weights is None:
            self._weights = self._default_weights() if self._default_weights is not None else {}
        return self._weights

    @weights.setter
    def weights(self, w: Dict[str, np.ndarray]) -> None:
        self._weights = w

    # -- weights ------------------------------------------------------------------------------
    def get_weights_dict(self) -> Dict[str, np.ndarray]:
        return self.weights

    def set_weights_dict(self, w: Dict[str, np.ndarray]) -> None:
        self.weights = w
        self._drop_plan()

    def _drop_plan(self) -> None:
        """Forget the plan and everything that holds raw pointers into it (captured CUDA graphs replay the packed-weight and
        scratch addresses they were captured with; the plan's destructor frees the former)."""
        for name in ("_pipe_graphs", "_phase_graphs", "_phase_cache", "_stream_cache", "_stream_ws"):
            self.__dict__.pop(name, None)
        self._plan = None

    def _workspace(self, plan: "Plan", batch: int, h: int = 0, w: int = 0) -> torch.Tensor:
        if self._ws_override is not None:
            need = plan.workspace_bytes(batch, h, w)
            if self._ws_override.numel() < need:
                raise RuntimeError(f"private scratch of {self._ws_override.numel()} bytes is smaller than the {need} this call needs")
            return self._ws_override
        return plan.workspace(batch, h, w)

    def count_params(self) -> int:
        return int(sum(int(np.prod(v.shape)) for v in self._flat_weights().values()))

    def _flat_weights(self) -> Dict[str, np.ndarray]:
        return self.weights

    # -- plan ---------------------------------------------------------------------------------
    def _make_plan(self, precision: str) -> Plan:
        raise NotImplementedError

    def plan(self) -> Plan:
        prec = runtime.get_precision()
        if self._plan is None or self._plan_precision != prec:
            self._drop_plan()
            self._plan = self._make_plan(prec)
            self._plan_precision = prec
            self._plan_gen += 1
        return self._plan

    # -- protocols ----------------------------------------------------------------------------
    def forward_device(self, inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        raise NotImplementedError

    def _run(self, inputs) -> List[torch.Tensor]:
        xs = [to_device_f32(x) for x in runtime.as_list(inputs)]
        return self.forward_device(xs)

    def predict(self, x, verbose=0, batch_size=None, reuse_output_buffers=False):
        """Keras-style predict: numpy (or torch, ideally pinned) in, numpy out.

        reuse_output_buffers=True returns views of per-model pinned staging buffers that the next
        predict() overwrites - the high-throughput mode (async D2H into pinned memory, one sync)."""
        outs = self._run(x)
        if not reuse_output_buffers:
            host = [o.cpu().numpy() for o in outs]
        else:
            stage = self.__dict__.setdefault("_stage", {})
            host_t = []
            for i, o in enumerate(outs):
                key = (i, tuple(o.shape), o.dtype)
                buf = stage.get(key)
                if buf is None:
                    buf = stage[key] = torch.empty(o.shape, dtype=o.dtype, pin_memory=True)
                buf.copy_(o, non_blocking=True)
                host_t.append(buf)
            torch.cuda.current_stream().synchronize()
            host = [b.numpy() for b in host_t]
        return host if self._multi_output else host[0]

    def predict_pipelined(self, x, n_chunks=4, on_chunk=None):
        """High-throughput predict for host inputs: the batch is cut into `n_chunks` contiguous chunks that flow
        through three CUDA streams - host->device copies, the model, device->host copies into per-model pinned
        staging buffers - so PCIe traffic overlaps the kernels.  Batch items are independent in every model of the
        path, so the result equals predict().  `on_chunk(device_inputs, device_outputs)` (optional) runs on the
        compute stream after each chunk (e.g. the metric kernels) and its return values are collected.
        Returns (host outputs - views of staging buffers the next call overwrites -, [on_chunk results])."""
        xs = runtime.as_list(x)
        hs = [t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)) for t in xs]
        hs = [t if t.dtype == torch.float32 else t.to(torch.float32) for t in hs]
        n = hs[0].shape[0]
        dev = runtime.require_cuda()
        if isinstance(n_chunks, (list, tuple)):                     # explicit chunk sizes (small first / last chunks shorten
            sizes = [int(v) for v in n_chunks if int(v) > 0]        # the exposed first upload and last download)
            if sum(sizes) != n:
                raise ValueError(f"chunk sizes {sizes} do not add up to the batch size {n}")
            bounds = [0]
            for v in sizes:
                bounds.append(bounds[-1] + v)
            n_chunks = len(sizes)
        else:
            n_chunks = max(1, min(int(n_chunks), n))
            bounds = [(i * n) // n_chunks for i in range(n_chunks + 1)]
        compute = torch.cuda.current_stream()
        st = self.__dict__.setdefault("_pipe_streams", {})
        if "in" not in st:
            st["in"], st["out"] = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        s_in, s_out = st["in"], st["out"]
        s_in.wait_stream(compute)
        s_out.wait_stream(compute)
        stage = self.__dict__.setdefault("_stage", {})
        keep, host, extra = [], None, []
        timeline = runtime.pipe_timeline()                               # debug: per-chunk event times of the three streams
        ev_in = [torch.cuda.Event(enable_timing=timeline) for _ in range(n_chunks)]
        if timeline:
            ev_t0 = torch.cuda.Event(enable_timing=True)
            ev_t0.record(compute)
            ev_cs, ev_os = [], []
        # Repeated calls with the same chunking replay one CUDA graph per chunk (the ~55 launches of a small chunk are
        # otherwise launch-bound on the host): the first call runs eagerly, then the chunks are captured with
        # persistent device input / output buffers.  runtime.set_cuda_graphs(False) keeps the eager path.
        self.plan()                                                 # (re)build first: a rebuild drops the graph caches
        gkey = (tuple(bounds), tuple(tuple(h.shape[1:]) for h in hs), id(on_chunk), self._plan_gen)
        gstate = self.__dict__.setdefault("_pipe_graphs", {})
        use_graphs = runtime.use_cuda_graphs()
        gs = gstate.get(gkey) if use_graphs else None
        if gs is not None and gs.get("calls", 0) >= 1 and "graphs" not in gs and not gs.get("failed"):
            try:
                gs.update(self._capture_pipe_graphs(bounds, hs, on_chunk, dev))
            except Exception as e:  # noqa: BLE001  (a path that synchronises cannot be captured: stay eager, say so once)
                gs["failed"] = True
                torch.cuda.synchronize()
                print(f"predict_pipelined: CUDA graph capture failed ({e!r}); using eager launches", file=sys.stderr)
        if gs is None and use_graphs:
            if len(gstate) > 1:
                gstate.clear()
            gs = gstate[gkey] = {"calls": 0, "on_chunk": on_chunk}   # the strong reference keeps id(on_chunk) from being reused
        if gs is not None:
            gs["calls"] += 1
        graphs = gs.get("graphs") if gs is not None else None
        for i in range(n_chunks):                                   # all uploads are queued up front on the copy-in stream
            lo, hi = bounds[i], bounds[i + 1]
            with torch.cuda.stream(s_in):
                if graphs is not None:
                    d_in = gs["d_in"][i]
                    for k, h in enumerate(hs):
                        d_in[k].copy_(h[lo:hi], non_blocking=True)
                else:
                    d_in = [h[lo:hi].to(dev, non_blocking=True) for h in hs]
                ev_in[i].record(s_in)
            keep.append(d_in)
        offs = None
        for i in range(n_chunks):
            lo, hi = bounds[i], bounds[i + 1]
            compute.wait_event(ev_in[i])
            if graphs is not None:
                graphs[i].replay()
                outs = gs["outs"][i]
                if on_chunk is not None:
                    extra.append(gs["extra"][i])
            else:
                outs = self.forward_device(keep[i])
                if on_chunk is not None:
                    extra.append(on_chunk(keep[i], getattr(self, "last", None)))
            ev_c = torch.cuda.Event(enable_timing=timeline)
            ev_c.record(compute)
            if host is None:                                        # staging buffers sized from the first chunk's row ratio
                per = [o.shape[0] // (hi - lo) for o in outs]
                host, offs = [], [0] * len(outs)
                for k, o in enumerate(outs):
                    shape = (per[k] * n,) + tuple(o.shape[1:])
                    key = ("pipe", k, shape, o.dtype)
                    buf = stage.get(key)
                    if buf is None:
                        buf = stage[key] = torch.empty(shape, dtype=o.dtype, pin_memory=True)
                    host.append(buf)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                for k, o in enumerate(outs):
                    host[k][offs[k]:offs[k] + o.shape[0]].copy_(o, non_blocking=True)
                    offs[k] += o.shape[0]
                if timeline:
                    ev_o = torch.cuda.Event(enable_timing=True)
                    ev_o.record(s_out)
                    ev_cs.append(ev_c)
                    ev_os.append(ev_o)
            keep.append(outs)                                       # alive until the copies have completed
        compute.wait_stream(s_out)
        s_out.synchronize()
        if timeline:
            torch.cuda.synchronize()
            print("pipe timeline (ms since start): " + "  ".join(
                f"[{bounds[i + 1] - bounds[i]}: in {ev_t0.elapsed_time(ev_in[i]):.2f} compute {ev_t0.elapsed_time(ev_cs[i]):.2f} out {ev_t0.elapsed_time(ev_os[i]):.2f}]"
                for i in range(n_chunks)), file=sys.stderr)
        res = [b.numpy() for b in host]
        return (res if self._multi_output else res[0]), extra

    def _capture_pipe_graphs(self, bounds, hs, on_chunk, dev):
        """One CUDA graph per chunk of predict_pipelined: forward (+ on_chunk) on persistent buffers."""
        d_ins, graphs, outs_all, extras = [], [], [], []
        torch.cuda.synchronize()
        plan = self.plan()
        need = 0
        for i in range(len(bounds) - 1):
            shp = hs[0].shape
            need = max(need, plan.workspace_bytes(bounds[i + 1] - bounds[i], *(shp[1:3] if len(shp) == 4 else (0, 0))))
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev)   # private scratch: the graphs bake its address in
        self._ws_override = ws
        try:
            for i in range(len(bounds) - 1):
                lo, hi = bounds[i], bounds[i + 1]
                d_in = [torch.zeros((hi - lo,) + tuple(h.shape[1:]), dtype=torch.float32, device=dev) for h in hs]
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    outs = self.forward_device(d_in)
                    ex = on_chunk(d_in, getattr(self, "last", None)) if on_chunk is not None else None
                d_ins.append(d_in)
                graphs.append(g)
                outs_all.append(outs)
                extras.append(ex)
        finally:
            self._ws_override = None
        torch.cuda.synchronize()
        return {"d_in": d_ins, "graphs": graphs, "outs": outs_all, "extra": extras, "ws": ws}

    def __call__(self, inputs, training=False):
        if training:
            raise NotImplementedError("training is outside the accelerated inference path (SURVEY.md §2 #19-20)")
        outs = [o.cpu() for o in self._run(inputs)]
        return outs if self._multi_output else outs[0]

    _multi_output = False

    def summary(self):
        print(f"Model: {self.name}  params: {self.count_params():,}")
        for k, v in self._flat_weights().items():
            print(f"  {k:40s} {tuple(v.shape)}")


# --------------------------------------------------------------------------------------------
class AutoencoderModel(Model):
    """`build_autoencoder(input_shape)` (train_autoencoder.py:9-40); shape-generic in H, W (% 4 == 0)."""

    name = "autoencoder"

    def __init__(self, input_shape, seed: int = 0):
        super().__init__()
        self.input_shape = tuple(int(v) for v in input_shape)
        self._default_weights = lambda: W.synthetic_autoencoder(seed=seed, channels=self.input_shape[2], keras_default=True)

    def _make_plan(self, precision):
        return Plan(_lib.PLAN_AUTOENCODER, self.weights, img_shape=(0, 0, self.input_shape[2]), precision=precision)

    def forward_device(self, inputs, want_u8: bool = False):
        x = inputs[0]
        if x.dim() != 4 or x.shape[3] != self.input_shape[2]:
            raise ValueError(f"expected (B,H,W,{self.input_shape[2]}) input, got {tuple(x.shape)}")
        b, h, w, c = x.shape
        if h % 4 or w % 4:
            raise ValueError(f"autoencoder needs H and W divisible by 4, got {h}x{w}")
        plan = self.plan()
        y = torch.empty_like(x)
        y8 = torch.empty((b, h, w, c), dtype=torch.uint8, device=x.device) if want_u8 else None
        ws = self._workspace(plan, b, h, w)
        _lib.check(_lib.lib.cic_autoencoder_forward(plan.handle, ptr(x), ptr(y), ptr(y8), b, h, w, ptr(ws), ws.numel(),
                                                    runtime.stream_ptr()))
        return [y, y8] if want_u8 else [y]


class EncoderModel(Model):
    """`build_encoder(img_shape, latent_dim, name, add_attention)` (GAN_functions.py:280-331)."""

    _multi_output = True

    def __init__(self, img_shape, latent_dim, name="encoder", add_attention=True, seed: int = 0):
        super().__init__()
        self.img_shape = tuple(int(v) for v in img_shape)
        self.latent_dim = int(latent_dim)
        self.name = name
        self.add_attention = bool(add_attention)
        self._default_weights = lambda: W.synthetic_encoder(self.img_shape, self.latent_dim, self.add_attention, seed, keras_default=True)

    def _make_plan(self, precision):
        return Plan(_lib.PLAN_ENCODER, self.weights, self.img_shape, self.latent_dim, self.add_attention, precision)

    def forward_device(self, inputs):
        x = inputs[0]
        h, w, c = self.img_shape
        if tuple(x.shape[1:]) != (h, w, c):
            raise ValueError(f"{self.name}: expected (B,{h},{w},{c}), got {tuple(x.shape)}")
        b = x.shape[0]
        dev = x.device
        lat = torch.empty((b, self.latent_dim), dtype=torch.float32, device=dev)
        x1 = torch.empty((b, h // 2, w // 2, 64), dtype=torch.float32, device=dev)
        x2 = torch.empty((b, h // 4, w // 4, 128), dtype=torch.float32, device=dev)
        x3 = torch.empty((b, h // 8, w // 8, 256), dtype=torch.float32, device=dev)
        plan = self.plan()
        ws = self._workspace(plan, b)
        _lib.check(_lib.lib.cic_encoder_forward(plan.handle, ptr(x), ptr(lat), ptr(x1), ptr(x2), ptr(x3), b, ptr(ws),
                                                ws.numel(), runtime.stream_ptr()))
        return [lat, x1, x2, x3]


class GeneratorModel(Model):
    """`build_generator(latent_dim, img_shape, name)` (GAN_functions.py:236-278)."""

    def __init__(self, latent_dim, img_shape, name="generator", seed: int = 0):
        super().__init__()
        self.img_shape = tuple(int(v) for v in img_shape)
        self.latent_dim = int(latent_dim)
        self.name = name
        self._default_weights = lambda: W.synthetic_generator(self.latent_dim, self.img_shape, seed, keras_default=True)

    def _make_plan(self, precision):
        return Plan(_lib.PLAN_GENERATOR, self.weights, self.img_shape, self.latent_dim, False, precision)

    def forward_device(self, inputs):
        if len(inputs) != 4:
            raise ValueError(f"{self.name} takes [latent, skip1, skip2, skip3]")
        lat, s1, s2, s3 = inputs
        h, w, c = self.img_shape
        b = lat.shape[0]
        want = [(b, self.latent_dim), (b, h // 2, w // 2, 64), (b, h // 4, w // 4, 128), (b, h // 8, w // 8, 256)]
        for t, s in zip(inputs, want):
            if tuple(t.shape) != s:
                raise ValueError(f"{self.name}: expected input shapes {want}, got {runtime.shapes_str(inputs)}")
        out = torch.empty((b, h, w, c), dtype=torch.float32, device=lat.device)
        plan = self.plan()
        ws = self._workspace(plan, b)
        _lib.check(_lib.lib.cic_generator_forward(plan.handle, ptr(lat), ptr(s1), ptr(s2), ptr(s3), ptr(out), b, ptr(ws),
                                                  ws.numel(), runtime.stream_ptr()))
        return [out]


class LatentSaliencyModel(Model):
    """`build_latent_saliency_model(latent_dim, name)` (GAN_functions.py:210-234)."""

    def __init__(self, latent_dim, name="latent_saliency_module", seed: int = 0):
        super().__init__()
        self.latent_dim = int(latent_dim)
        self.name = name
        self._default_weights = lambda: W.synthetic_latent_saliency(self.latent_dim, seed, keras_default=True)

    def _make_plan(self, precision):
        return Plan(_lib.PLAN_SALIENCY, self.weights, (0, 0, 3), self.latent_dim, False, precision)

    def forward_device(self, inputs):
        lat = inputs[0]
        if lat.dim() != 2 or lat.shape[1] != self.latent_dim:
            raise ValueError(f"{self.name}: expected (B,{self.latent_dim}), got {tuple(lat.shape)}")
        b = lat.shape[0]
        out = torch.empty((b, 1), dtype=torch.float32, device=lat.device)
        plan = self.plan()
        ws = self._workspace(plan, b)
        _lib.check(_lib.lib.cic_saliency_forward(plan.handle, ptr(lat), ptr(out), b, ptr(ws), ws.numel(), runtime.stream_ptr()))
        return [out]


class RDOptimizerModel(Model):
    """`build_rate_distortion_optimizer(img_shape, latent_dims, name)` (GAN_functions.py:495-557).

    Inputs `[img, saliency, target_bpp]` like the reference; the image is unused by the graph (:500).
    """

    def __init__(self, img_shape, latent_dims=None, name="rd_optimizer", seed: int = 0):
        super().__init__()
        self.img_shape = tuple(int(v) for v in img_shape)
        self.latent_dims = latent_dims
        self.name = name
        self._default_weights = lambda: W.synthetic_rd_optimizer(seed, keras_default=True)

    def _make_plan(self, precision):
        return Plan(_lib.PLAN_RD, self.weights, self.img_shape, 0, False, precision)

    def forward_device(self, inputs):
        if len(inputs) == 3:
            _, mask, bpp = inputs
        elif len(inputs) == 2:
            mask, bpp = inputs
        else:
            raise ValueError(f"{self.name} takes [img, saliency, target_bpp]")
        h, w, _ = self.img_shape
        b = mask.shape[0]
        if tuple(mask.shape) != (b, h, w, 1):
            raise ValueError(f"{self.name}: saliency must be (B,{h},{w},1), got {tuple(mask.shape)}")
        bpp = bpp.reshape(-1).contiguous()
        out = torch.empty((b, 3), dtype=torch.float32, device=mask.device)
        plan = self.plan()
        ws = self._workspace(plan, b)
        _lib.check(_lib.lib.cic_rd_forward(plan.handle, ptr(mask), ptr(bpp), ptr(out), b, ptr(ws), ws.numel(), runtime.stream_ptr()))
        return [out]


class AdaptiveCompressionModel(Model):
    """The 3-input / 5-output `adaptive_model` of build_adaptive_compression_model (GAN_functions.py:686-698).

    predict([img, mask, bpp]) -> [blended, hq_latent_quantized, lq_latent_quantized, rd_params, dynamic_threshold].
    Images may be any multiple of the model's tile (256 in the reference); they are coded as independent tiles,
    and the latent outputs then have one row per tile in (image, tile-row, tile-col) order.
    """

    name = "adaptive_compression_model"
    _multi_output = True
    SUBS = ("hq_encoder", "hq_generator", "lq_encoder", "lq_generator", "latent_saliency_hq", "latent_saliency_lq",
            "rd_optimizer")

    def __init__(self, img_shape, base_latent_dim, components: Dict[str, Model]):
        super().__init__()
        self.img_shape = tuple(int(v) for v in img_shape)
        self.base_latent_dim = int(base_latent_dim)
        self.components = components
        self._seen = None

    def _flat_weights(self):
        flat = {}
        for sub in self.SUBS:
            for k, v in self.components[sub].weights.items():
                flat[f"{sub}/{k}"] = v
        return flat

    def set_weights_dict(self, w):
        """Accepts {sub_model: {name: array}} (weights.synthetic_adaptive) and shares it with the sub-models."""
        for sub in self.SUBS:
            self.components[sub].set_weights_dict(w[sub])
        self._drop_plan()

    def get_weights_dict(self):
        return {sub: self.components[sub].weights for sub in self.SUBS}

    def plan(self):
        # sub-models may have had their weights replaced individually: rebuild when identities change
        ids = tuple(id(self.components[s].weights) for s in self.SUBS)
        if self._seen != ids:
            self._drop_plan()
            self._seen = ids
        return super().plan()

    def _make_plan(self, precision):
        return Plan(_lib.PLAN_ADAPTIVE, self._flat_weights(), self.img_shape, self.base_latent_dim, False, precision)

    # ---- geometry ------------------------------------------------------------------------------------------------------------
    PADDING = "edge"   # how images that are not a multiple of the model tile are extended: the last row / column
                       # of the image is replicated into the ragged tiles on load, outputs are cropped on store

    def tiles_per_image(self, h: int, w: int) -> int:
        T = self.img_shape[0]
        return (-(-h // T)) * (-(-w // T))

    def _check_inputs(self, img, mask, bpp):
        n, h, w, c = img.shape
        if c != 3 or h < 1 or w < 1:
            raise ValueError(f"image must be (B, H, W, 3), got {tuple(img.shape)}")
        if mask.dim() == 3:
            mask = mask.unsqueeze(-1)
        if tuple(mask.shape) != (n, h, w, 1):
            raise ValueError(f"saliency must be ({n},{h},{w},1), got {tuple(mask.shape)}")
        bpp = bpp.reshape(-1)
        if bpp.numel() != n:
            raise ValueError(f"target_bpp must have one value per image ({n}), got {bpp.numel()}")
        return mask, bpp

    def forward_device(self, inputs, extras: bool = False):
        """[image (n,H,W,3), saliency (n,H,W,1), target_bpp (n,)] on the device -> the five model outputs on the device.
        H, W of any size: the image is coded as ceil(H/T) x ceil(W/T) independent tiles of the model size T (256 in the
        reference, whose graph is fixed at that size: GAN_functions.py:242-248); ragged tiles replicate the image edge and are
        cropped, dt / blend / hq_ratio cover the H x W pixels of the image only."""
        if len(inputs) != 3:
            raise ValueError("adaptive model takes [image, saliency, target_bpp]")
        img, mask, bpp = inputs
        if img.dim() != 4:
            raise ValueError(f"image must be (B, H, W, 3), got {tuple(img.shape)}")
        mask, bpp = self._check_inputs(img, mask, bpp)
        bpp = bpp.contiguous()
        n, h, w, c = img.shape
        nt = n * self.tiles_per_image(h, w)
        dev = img.device
        base = self.base_latent_dim
        f32 = dict(dtype=torch.float32, device=dev)
        out = {
            "blended": torch.empty((n, h, w, 3), **f32),
            "hq_latent_q": torch.empty((nt, 2 * base), **f32),
            "lq_latent_q": torch.empty((nt, base), **f32),
            "rd_params": torch.empty((nt, 3), **f32),
            "dt": torch.empty((n, h, w, 1), **f32),
            "hq_ratio_sum": torch.empty((n,), dtype=torch.float64, device=dev),
        }
        if extras:
            out.update({
                "hq_symbols": torch.empty((nt, 2 * base), dtype=torch.int32, device=dev),
                "lq_symbols": torch.empty((nt, base), dtype=torch.int32, device=dev),
                "hq_latent": torch.empty((nt, 2 * base), **f32),
                "lq_latent": torch.empty((nt, base), **f32),
                "hq_scale": torch.empty((nt,), **f32),
                "lq_scale": torch.empty((nt,), **f32),
                "hq_out": torch.empty((n, h, w, 3), **f32),
                "lq_out": torch.empty((n, h, w, 3), **f32),
            })
        io = _lib.cic_adaptive_io()
        io.d_img, io.d_mask, io.d_bpp = ptr(img), ptr(mask), ptr(bpp)
        io.d_blended, io.d_hq_latent_q, io.d_lq_latent_q = ptr(out["blended"]), ptr(out["hq_latent_q"]), ptr(out["lq_latent_q"])
        io.d_rd_params, io.d_dt, io.d_hq_ratio_sum = ptr(out["rd_params"]), ptr(out["dt"]), ptr(out["hq_ratio_sum"])
        if extras:
            io.d_hq_symbols, io.d_lq_symbols = ptr(out["hq_symbols"]), ptr(out["lq_symbols"])
            io.d_hq_latent, io.d_lq_latent = ptr(out["hq_latent"]), ptr(out["lq_latent"])
            io.d_hq_scale, io.d_lq_scale = ptr(out["hq_scale"]), ptr(out["lq_scale"])
            io.d_hq_out, io.d_lq_out = ptr(out["hq_out"]), ptr(out["lq_out"])
        plan = self.plan()
        ws = self._workspace(plan, n, h, w)
        _lib.check(_lib.lib.cic_adaptive_forward(plan.handle, C.byref(io), n, h, w, ptr(ws), ws.numel(), runtime.stream_ptr()))
        self.last = out
        self._last_inputs = [img, mask, bpp]
        if extras:
            return out
        return [out["blended"], out["hq_latent_q"], out["lq_latent_q"], out["rd_params"], out["dt"]]

    # ---- phased, pipelined predict -------------------------------------------------------------------------------------------
    def _phase_buffers(self, n, h, w, dev):
        """Persistent device buffers of predict_phased for one batch geometry: inputs, outputs and the batch-wide state."""
        self.plan()
        key = (n, h, w, self._plan_gen)
        cache = self.__dict__.setdefault("_phase_cache", {})
        if key in cache:
            return cache[key]
        if len(cache) > 1:                       # the graphs captured on the evicted buffers go with them
            cache.clear()
            self.__dict__.pop("_phase_graphs", None)
        T, base = self.img_shape[0], self.base_latent_dim
        tpi = self.tiles_per_image(h, w)
        nt = n * tpi
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        b = {
            "img": torch.empty((n, h, w, 3), **f32), "mask": torch.empty((n, h, w, 1), **f32), "bpp": torch.empty((n,), **f32),
            "blended": torch.empty((n, h, w, 3), **f32), "dt": torch.empty((n, h, w, 1), **f32),
            "hq_latent_q": torch.empty((nt, 2 * base), **f32), "lq_latent_q": torch.empty((nt, base), **f32),
            "rd_params": torch.empty((nt, 3), **f32), "hq_ratio_sum": torch.empty((n,), dtype=torch.float64, device=dev),
        }
        e = [(T // 2) ** 2 * 64, (T // 4) ** 2 * 128, (T // 8) ** 2 * 256, (T // 16) ** 2 * 512]
        st = _lib.cic_adaptive_state()
        keep = []
        for name, elems in (("x1", e[0]), ("x2", e[1]), ("x3", e[2]), ("x4_hi", e[3]), ("x4_lo", e[3]), ("g0", e[3])):
            for k in range(2):
                t = torch.empty((nt, elems), **bf)
                keep.append(t)
                getattr(st, name)[k] = t.data_ptr()
        b["img_u8"] = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        b["blended_u8"] = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        b["state"], b["state_keep"], b["tpi"], b["nt"] = st, keep, tpi, nt
        cache[key] = b
        return b

    def _phase_call(self, b, phase, lo, hi, h, w, want_dt=True):
        """One cic_adaptive_forward_phase call on images [lo, hi) of the persistent buffers (LATENT: the whole batch)."""
        io = _lib.cic_adaptive_io()
        tpi = b["tpi"]
        if phase == _lib.PHASE_LATENT:
            io.d_bpp = ptr(b["bpp"])
            io.d_hq_latent_q, io.d_lq_latent_q = ptr(b["hq_latent_q"]), ptr(b["lq_latent_q"])
            lo, hi, tile0 = 0, b["bpp"].shape[0], 0
        else:
            tile0 = lo * tpi
            io.d_img, io.d_mask, io.d_bpp = ptr(b["img"][lo:hi]), ptr(b["mask"][lo:hi]), ptr(b["bpp"][lo:hi])
            if phase == _lib.PHASE_ENCODE:
                io.d_rd_params = ptr(b["rd_params"][tile0:hi * tpi])
            else:
                io.d_blended = ptr(b["blended"][lo:hi])
                io.d_dt = ptr(b["dt"][lo:hi]) if want_dt else None
                io.d_hq_ratio_sum = ptr(b["hq_ratio_sum"][lo:hi])
        plan = self.plan()
        ws = self._workspace(plan, hi - lo, h, w)
        _lib.check(_lib.lib.cic_adaptive_forward_phase(plan.handle, C.byref(io), C.byref(b["state"]), phase, tile0, hi - lo, h, w,
                                                       ptr(ws), ws.numel(), runtime.stream_ptr()))

    def rate_sweep_device(self, img: torch.Tensor, mask: torch.Tensor, levels, on_level=None):
        """The rate-control sweep of GAN_test.py:532-645 WITH reconstructions (BASELINE configs[2]): the reference runs the whole
        model once per target bpp; the encoder convolutions and skips do not depend on the target, so they run once here and only
        the Dense / latent-saliency / quantiser phase and the decoders + ROI blend run per level - same results as one full
        predict per level.  img (n,H,W,3), mask (n,H,W,1) on the device; levels: target bpps.  `on_level(k, inputs, outputs)`
        is called after every level with views of the persistent buffers (valid until the next level).  Returns hq_ratio
        (levels, n) float64 on the device."""
        if mask.dim() == 3:
            mask = mask.unsqueeze(-1)
        n, h, w, _ = img.shape
        dev = img.device
        b = self._phase_buffers(n, h, w, dev)
        b["img"].copy_(img)
        b["mask"].copy_(mask)
        levels = [float(v) for v in levels]
        ratios = torch.empty((len(levels), n), dtype=torch.float64, device=dev)
        b["bpp"].fill_(levels[0] if levels else 1.0)
        self._phase_call(b, _lib.PHASE_ENCODE, 0, n, h, w)
        for k, level in enumerate(levels):
            b["bpp"].fill_(level)
            self._phase_call(b, _lib.PHASE_LATENT, 0, n, h, w)
            self._phase_call(b, _lib.PHASE_DECODE, 0, n, h, w)
            ratios[k].copy_(b["hq_ratio_sum"])
            if on_level is not None:
                on_level(k, [b["img"], b["mask"], b["bpp"]],
                         {"blended": b["blended"], "dt": b["dt"], "hq_ratio_sum": b["hq_ratio_sum"],
                          "hq_latent_q": b["hq_latent_q"], "lq_latent_q": b["lq_latent_q"]})
        return ratios / float(h * w)

    # ---- streaming predict: batches overlap each other ----------------------------------------------------------------------
    def _stream_slot(self, slot, n, h, w, dev, u8_io, want_dt, jpeg_cap=0):
        """Persistent device + pinned host buffers of one slot of predict_stream."""
        self.plan()
        key = ("stream", slot, n, h, w, self._plan_gen, bool(u8_io), bool(want_dt), int(jpeg_cap))
        cache = self.__dict__.setdefault("_stream_cache", {})
        if key in cache:
            return cache[key]
        for k in [k for k in cache if k[1] == slot]:            # another geometry used this slot: drop it (and its graph)
            del cache[k]
        base = self.base_latent_dim
        nt = n * self.tiles_per_image(h, w)
        f32 = dict(dtype=torch.float32, device=dev)
        b = {
            "img": torch.empty((n, h, w, 3), **f32), "mask": torch.empty((n, h, w, 1), **f32), "bpp": torch.empty((n,), **f32),
            "blended": torch.empty((n, h, w, 3), **f32), "dt": torch.empty((n, h, w, 1), **f32) if want_dt else None,
            "hq_latent_q": torch.empty((nt, 2 * base), **f32), "lq_latent_q": torch.empty((nt, base), **f32),
            "rd_params": torch.empty((nt, 3), **f32), "hq_ratio_sum": torch.empty((n,), dtype=torch.float64, device=dev),
        }
        if u8_io:
            b["img_u8"] = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
            b["blended_u8"] = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        outs = [b["blended_u8"] if u8_io else b["blended"], b["hq_latent_q"], b["lq_latent_q"], b["rd_params"],
                b["dt"] if want_dt else b["hq_ratio_sum"]]
        if jpeg_cap:                                              # the reconstruction leaves the device as JPEG files
            b["jpeg"] = torch.empty((n, jpeg_cap), dtype=torch.uint8, device=dev)
            b["jpeg_sizes"] = torch.zeros((n,), dtype=torch.int32, device=dev)
            outs[0] = b["jpeg"]
            outs.append(b["jpeg_sizes"])
        b["outs"] = outs
        b["host"] = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]
        b["calls"], b["graph"], b["extra"] = 0, None, None
        cache[key] = b
        return b

    def _stream_compute(self, b, n, h, w, u8_io, want_dt, on_batch):
        """The whole forward of one batch on the slot's buffers (+ on_batch): what the slot's CUDA graph holds."""
        if u8_io:
            _lib.check(_lib.lib.cic_u8_to_f32_signed(ptr(b["img_u8"]), ptr(b["img"]), n * h * w * 3, runtime.stream_ptr()))
        io = _lib.cic_adaptive_io()
        io.d_img, io.d_mask, io.d_bpp = ptr(b["img"]), ptr(b["mask"]), ptr(b["bpp"])
        io.d_blended, io.d_hq_latent_q, io.d_lq_latent_q = ptr(b["blended"]), ptr(b["hq_latent_q"]), ptr(b["lq_latent_q"])
        io.d_rd_params, io.d_hq_ratio_sum = ptr(b["rd_params"]), ptr(b["hq_ratio_sum"])
        io.d_dt = ptr(b["dt"]) if want_dt else None
        plan = self.plan()
        ws = self._workspace(plan, n, h, w)
        _lib.check(_lib.lib.cic_adaptive_forward(plan.handle, C.byref(io), n, h, w, ptr(ws), ws.numel(), runtime.stream_ptr()))
        ex = None
        if on_batch is not None:
            ex = on_batch([b["img"], b["mask"], b["bpp"]], {"blended": b["blended"], "dt": b["dt"], "hq_ratio_sum": b["hq_ratio_sum"]})
        if u8_io:
            _lib.check(_lib.lib.cic_f32_signed_to_u8(ptr(b["blended"]), ptr(b["blended_u8"]), n * h * w * 3, runtime.stream_ptr()))
        if "jpeg" in b:                                           # save_image's cv2.imwrite on the device (RGB -> BGR folded in)
            jws = self.__dict__.setdefault("_stream_jpeg_ws", {})
            need = int(_lib.lib.cic_jpeg_workspace_bytes(n, h, w))
            if jws.get("bytes", 0) < need:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("JPEG workspace must exist before graph capture")
                for other in self.__dict__.get("_stream_cache", {}).values():      # graphs of other slots hold the old buffer
                    other["graph"] = None
                jws.update(bytes=need, buf=torch.empty(need, dtype=torch.uint8, device=b["jpeg"].device))
            _lib.check(_lib.lib.cic_jpeg_encode_u8(ptr(b["blended_u8"]), n, h, w, 1, 95, ptr(b["jpeg"]), b["jpeg"].shape[1], ptr(b["jpeg_sizes"]),
                                                   ptr(jws["buf"]), jws["bytes"], runtime.stream_ptr()))
        return ex

    def predict_stream(self, batches, on_batch=None, u8_io: bool = False, want_dt: bool = True, depth: int = 2, jpeg_out: float = 0.0):
        """predict() for a STREAM of host batches at full throughput: a generator that yields (host outputs, on_batch result) per
        batch, in order.  Batch k + 1 is uploaded while batch k is on the tensor cores and batch k - 1 travels back to pinned host
        memory (three CUDA streams, `depth` buffer sets), so in steady state a batch costs max(kernels, copies) - not their sum,
        and the kernels run on the whole batch (chunking a batch to hide its own copies, as predict_phased does for a single
        call, costs tile-quantisation efficiency on the small chunks).  From its second use a slot replays ONE CUDA graph
        (forward + on_batch).  Results equal predict().

        batches: iterable of [image (n,H,W,3), saliency (n,H,W,1), target_bpp (n,)] with one geometry; pinned tensors copy fastest.
        The yielded arrays are views of the slot's pinned buffers: valid until `depth` more batches have been yielded.
        u8_io / want_dt: as predict_phased (uint8 image up, uint8 reconstruction down; hq_ratio instead of the dt map).
        jpeg_out (with u8_io): bytes per pixel reserved for a JPEG file of every reconstruction (e.g. 1.0).  The reconstruction then
        leaves the device the way the reference stores it - save_image's cv2.imwrite(".jpg") (GAN_functions.py:41-50, GAN_test.py:390),
        encoded on the GPU, byte-identical to OpenCV's file: output 0 is a (n, capacity) uint8 array of files and a sixth output
        holds their sizes (file i = out[0][i, :out[5][i]]; a size above the capacity means that file was cut and needs a larger
        jpeg_out)."""
        dev = runtime.require_cuda()
        if jpeg_out and not u8_io:
            raise ValueError("jpeg_out needs u8_io=True (the files are made from the uint8 reconstruction)")
        compute = torch.cuda.current_stream()
        st = self.__dict__.setdefault("_pipe_streams", {})
        if "in" not in st:
            st["in"], st["out"] = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        s_in, s_out = st["in"], st["out"]
        s_in.wait_stream(compute)
        s_out.wait_stream(compute)
        depth = max(1, int(depth))
        pending = []                                                  # [(slot buffers, ev_out, extra, pixels per image)]

        def finish(item):
            b, ev_out, extra, hw = item
            ev_out.synchronize()
            res = [t.numpy() for t in b["host"]]
            if not want_dt:
                res[4] = res[4] / float(hw)
            return res, extra

        for k, x in enumerate(batches):
            if len(pending) == depth:
                yield finish(pending.pop(0))
            xs = runtime.as_list(x)
            hs = [t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)) for t in xs]
            if len(hs) != 3:
                raise ValueError("adaptive model takes [image, saliency, target_bpp]")
            if u8_io and hs[0].dtype != torch.uint8:
                raise ValueError("u8_io=True takes a uint8 image")
            hs = [t if (t.dtype == torch.float32 or (u8_io and i == 0)) else t.to(torch.float32) for i, t in enumerate(hs)]
            img, mask, bpp = hs
            if img.dim() != 4:
                raise ValueError(f"image must be (B, H, W, 3), got {tuple(img.shape)}")
            mask, bpp = self._check_inputs(img, mask, bpp)
            n, h, w, _ = img.shape
            jpeg_cap = (1024 + int(float(jpeg_out) * h * w) + 3) & ~3 if jpeg_out else 0
            b = self._stream_slot(k % depth, n, h, w, dev, u8_io, want_dt, jpeg_cap)
            timeline = runtime.pipe_timeline()
            ev_in, ev_c, ev_out = (torch.cuda.Event(enable_timing=timeline) for _ in range(3))
            if timeline:                                              # debug: start / end of the three legs of every batch
                tl = self.__dict__.setdefault("_stream_timeline", [])
                marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                marks[0].record(s_in)
            # the slot's previous batch (k - depth) has been yielded above, i.e. its download - hence its kernels - completed: the
            # slot's buffers are free, and this upload may overlap the kernels of batch k - 1 in the other slot
            with torch.cuda.stream(s_in):
                (b["img_u8"] if u8_io else b["img"]).copy_(img, non_blocking=True)
                b["mask"].copy_(mask, non_blocking=True)
                b["bpp"].copy_(bpp, non_blocking=True)
                ev_in.record(s_in)
            compute.wait_event(ev_in)
            if timeline:
                marks[1].record(compute)
            b["calls"] += 1
            if runtime.use_cuda_graphs() and b["calls"] >= 2 and b["graph"] is None and not b.get("failed") and b.get("on_batch_id") == id(on_batch):
                try:
                    torch.cuda.synchronize()
                    need = self.plan().workspace_bytes(n, h, w)
                    wsd = self.__dict__.setdefault("_stream_ws", {})
                    if wsd.get("gen") != self._plan_gen or wsd["buf"].numel() < need:   # one private scratch for all slots: their
                        wsd.clear()                                                    # graphs run one after the other on one stream
                        for other in self.__dict__.get("_stream_cache", {}).values():
                            other["graph"] = None
                        wsd.update(gen=self._plan_gen, buf=torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev))
                    self._ws_override = wsd["buf"]
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        ex = self._stream_compute(b, n, h, w, u8_io, want_dt, on_batch)
                    b["graph"], b["extra"], b["on_batch"] = g, ex, on_batch
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    b["failed"] = True
                    torch.cuda.synchronize()
                    print(f"predict_stream: CUDA graph capture failed ({e!r}); using eager launches", file=sys.stderr)
                finally:
                    self._ws_override = None
            if b["graph"] is not None and b.get("on_batch") is on_batch:
                b["graph"].replay()
                extra = b["extra"]
            else:
                b["on_batch_id"] = id(on_batch)
                b["on_batch_keep"] = on_batch
                extra = self._stream_compute(b, n, h, w, u8_io, want_dt, on_batch)
            ev_c.record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                if timeline:
                    marks[2].record(s_out)
                for hbuf, o in zip(b["host"], b["outs"]):
                    hbuf.copy_(o, non_blocking=True)
                ev_out.record(s_out)
            if timeline:
                tl.append((marks[0], ev_in, marks[1], ev_c, marks[2], ev_out))
            pending.append((b, ev_out, extra, h * w))
        while pending:
            yield finish(pending.pop(0))
        compute.wait_stream(s_out)
        if runtime.pipe_timeline() and self.__dict__.get("_stream_timeline"):
            torch.cuda.synchronize()
            tl = self.__dict__.pop("_stream_timeline")
            t0 = tl[0][0]
            for i, ev in enumerate(tl):
                t = [t0.elapsed_time(e) for e in ev]
                print(f"stream batch {i}: upload {t[0]:.2f}-{t[1]:.2f}  kernels {t[2]:.2f}-{t[3]:.2f}  download {t[4]:.2f}-{t[5]:.2f} ms", flush=True)

    def predict_phased(self, x, enc_chunks=None, dec_chunks=None, on_chunk=None, u8_io: bool = False, want_dt: bool = True):
        """predict() for host batches at full throughput: the forward is cut into three phases (include/cic.h): the encoder
        convolutions run per upload chunk while the next chunk's host->device copy is in flight, the Dense / saliency /
        quantiser phase runs once on the whole batch (its 1.2 GB of Dense weights are streamed once instead of once per
        chunk), and the decoders run per download chunk while the previous chunk's outputs travel to pinned host buffers.
        Same results as predict().  `on_chunk(device_inputs, outputs)` runs on the compute stream after every decode chunk.
        Returns (host outputs, [on_chunk results]); the host outputs are views of pinned buffers the next call overwrites.

        u8_io=True: the reference's file-boundary pixel format on the wire.  The image is uint8 RGB and is normalised on the
        device ((u8 - 127.5) / 127.5: load_and_preprocess_image, GAN_functions.py:31-37), the blended output comes back as uint8
        (((x + 1) * 127.5).astype(uint8): save_image, :41-50): 1 instead of 4 bytes per sample over PCIe.  on_chunk still sees
        the float32 tensors, so PSNR / SSIM are the reference's (computed before the uint8 cast, GAN_test.py:297-300).
        want_dt=False: the bit-allocation map is not downloaded; the fifth host output is hq_ratio = mean(dt) per image
        (float64, (n,)) - all GAN_test.py:312,573 take from it."""
        xs = runtime.as_list(x)
        hs = [t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)) for t in xs]
        if len(hs) != 3:
            raise ValueError("adaptive model takes [image, saliency, target_bpp]")
        if u8_io and hs[0].dtype != torch.uint8:
            raise ValueError("u8_io=True takes a uint8 image")
        hs = [t if (t.dtype == torch.float32 or (u8_io and k == 0)) else t.to(torch.float32) for k, t in enumerate(hs)]
        img, mask, bpp = hs
        if img.dim() != 4:
            raise ValueError(f"image must be (B, H, W, 3), got {tuple(img.shape)}")
        mask, bpp = self._check_inputs(img, mask, bpp)
        n, h, w, c = img.shape

        def bounds_of(sizes, default):
            sizes = [int(v) for v in (sizes if sizes is not None else default) if int(v) > 0]
            if sum(sizes) != n:
                raise ValueError(f"chunk sizes {sizes} do not add up to the batch size {n}")
            out = [0]
            for v in sizes:
                out.append(out[-1] + v)
            return out
        default = phased_default_chunks(n)
        eb = bounds_of(enc_chunks, default)
        db = bounds_of(dec_chunks, default[::-1])
        dev = runtime.require_cuda()
        b = self._phase_buffers(n, h, w, dev)
        plan = self.plan()
        compute = torch.cuda.current_stream()
        st = self.__dict__.setdefault("_pipe_streams", {})
        if "in" not in st:
            st["in"], st["out"] = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        s_in, s_out = st["in"], st["out"]
        s_in.wait_stream(compute)
        s_out.wait_stream(compute)
        stage = self.__dict__.setdefault("_stage", {})
        names = ["blended", "hq_latent_q", "lq_latent_q", "rd_params", "dt"]
        host = []
        for k in names:
            src = b["blended_u8"] if (u8_io and k == "blended") else (b["hq_ratio_sum"] if (k == "dt" and not want_dt) else b[k])
            key = ("phased", k, tuple(src.shape), src.dtype)
            if key not in stage:
                stage[key] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            host.append(stage[key])

        def encode(lo, hi):
            if u8_io:
                _lib.check(_lib.lib.cic_u8_to_f32_signed(ptr(b["img_u8"][lo:hi]), ptr(b["img"][lo:hi]), (hi - lo) * h * w * 3,
                                                         runtime.stream_ptr()))
            self._phase_call(b, _lib.PHASE_ENCODE, lo, hi, h, w)

        def decode(lo, hi):
            self._phase_call(b, _lib.PHASE_DECODE, lo, hi, h, w, want_dt)
            ex = None
            if on_chunk is not None:
                ex = on_chunk([b["img"][lo:hi], b["mask"][lo:hi], b["bpp"][lo:hi]],
                              {"blended": b["blended"][lo:hi], "dt": b["dt"][lo:hi] if want_dt else None,
                               "hq_ratio_sum": b["hq_ratio_sum"][lo:hi]})
            if u8_io:
                _lib.check(_lib.lib.cic_f32_signed_to_u8(ptr(b["blended"][lo:hi]), ptr(b["blended_u8"][lo:hi]), (hi - lo) * h * w * 3,
                                                         runtime.stream_ptr()))
            return ex
        # CUDA graphs: the first call with a chunking runs eagerly, the second captures one graph per phase call (persistent
        # buffers + a private scratch buffer make them replayable), later calls replay.  The cache entry owns everything whose
        # address the graphs bake in, and is keyed by the plan generation, so a rebuilt plan can never meet an old graph.
        gkey = (tuple(eb), tuple(db), n, h, w, id(on_chunk), self._plan_gen, bool(u8_io), bool(want_dt))
        gstate = self.__dict__.setdefault("_phase_graphs", {})
        gs = gstate.get(gkey)
        if gs is None:
            if len(gstate) > 1:
                gstate.clear()
            gs = gstate[gkey] = {"calls": 0, "on_chunk": on_chunk, "buffers": b}
        gs["calls"] += 1
        use_graphs = runtime.use_cuda_graphs()
        if use_graphs and gs["calls"] >= 2 and "enc" not in gs and not gs.get("failed"):
            try:
                torch.cuda.synchronize()
                sizes = [eb[i + 1] - eb[i] for i in range(len(eb) - 1)] + [db[i + 1] - db[i] for i in range(len(db) - 1)] + [n]
                need = max(plan.workspace_bytes(v, h, w) for v in set(sizes))
                gs["ws"] = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev)
                self._ws_override = gs["ws"]
                enc, dec, extra_g = [], [], []
                for i in range(len(eb) - 1):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        encode(eb[i], eb[i + 1])
                    enc.append(g)
                lat = torch.cuda.CUDAGraph()
                with torch.cuda.graph(lat, capture_error_mode="thread_local"):
                    self._phase_call(b, _lib.PHASE_LATENT, 0, n, h, w)
                for i in range(len(db) - 1):
                    lo, hi = db[i], db[i + 1]
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        ex = decode(lo, hi)
                    dec.append(g)
                    extra_g.append(ex)
                torch.cuda.synchronize()
                gs.update({"enc": enc, "lat": lat, "dec": dec, "extra": extra_g})
            except Exception as e:  # noqa: BLE001
                gs["failed"] = True
                gs.pop("ws", None)
                torch.cuda.synchronize()
                print(f"predict_phased: CUDA graph capture failed ({e!r}); using eager launches", file=sys.stderr)
            finally:
                self._ws_override = None
        graphs = gs if "enc" in gs else None
        # uploads: all queued up front on the copy-in stream
        timeline = runtime.pipe_timeline()                               # debug: event times of the three streams
        ev_in = [torch.cuda.Event(enable_timing=timeline) for _ in range(len(eb) - 1)]
        marks = []
        if timeline:
            ev_t0 = torch.cuda.Event(enable_timing=True)
            ev_t0.record(compute)

        def mark(name, stream):
            if timeline:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                marks.append((name, e))
        with torch.cuda.stream(s_in):
            b["bpp"].copy_(bpp, non_blocking=True)
            for i in range(len(eb) - 1):
                lo, hi = eb[i], eb[i + 1]
                (b["img_u8"] if u8_io else b["img"])[lo:hi].copy_(img[lo:hi], non_blocking=True)
                b["mask"][lo:hi].copy_(mask[lo:hi], non_blocking=True)
                ev_in[i].record(s_in)
        for i in range(len(eb) - 1):
            compute.wait_event(ev_in[i])
            if graphs is not None:
                graphs["enc"][i].replay()
            else:
                encode(eb[i], eb[i + 1])
            mark(f"enc{eb[i + 1] - eb[i]}", compute)
        if graphs is not None:
            graphs["lat"].replay()
        else:
            self._phase_call(b, _lib.PHASE_LATENT, 0, n, h, w)
        mark("latent", compute)
        ev_l = torch.cuda.Event()
        ev_l.record(compute)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_l)
            for k in (1, 2, 3):
                host[k].copy_(b[names[k]], non_blocking=True)
        extra = []
        for i in range(len(db) - 1):
            lo, hi = db[i], db[i + 1]
            if graphs is not None:
                graphs["dec"][i].replay()
                if on_chunk is not None:
                    extra.append(graphs["extra"][i])
            else:
                ex = decode(lo, hi)
                if on_chunk is not None:
                    extra.append(ex)
            mark(f"dec{hi - lo}", compute)
            ev_c = torch.cuda.Event()
            ev_c.record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                host[0][lo:hi].copy_((b["blended_u8"] if u8_io else b["blended"])[lo:hi], non_blocking=True)
                host[4][lo:hi].copy_((b["dt"] if want_dt else b["hq_ratio_sum"])[lo:hi], non_blocking=True)
                mark(f"out{hi - lo}", s_out)
        compute.wait_stream(s_out)
        s_out.synchronize()
        if timeline:
            torch.cuda.synchronize()
            print("phase timeline (ms since start): in " + " ".join(f"{ev_t0.elapsed_time(e):.2f}" for e in ev_in) + " | " +
                  "  ".join(f"{nm} {ev_t0.elapsed_time(e):.2f}" for nm, e in marks), file=sys.stderr)
        res = [t.numpy() for t in host]
        if not want_dt:
            res[4] = res[4] / float(h * w)
        return res, extra
