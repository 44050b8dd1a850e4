"""ctypes binding of libcic.so (the C ABI declared in include/cic.h).

There is no CPU fallback: if the shared library has not been built this module raises at import,
and every compute call raises `CicError` when the library reports a failure (e.g. no CUDA device).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcic.so")

# enums of include/cic.h
ACT_NONE, ACT_RELU, ACT_LRELU02, ACT_SIGMOID, ACT_TANH = range(5)
PREC_FP32, PREC_TC = 0, 1
PLAN_AUTOENCODER, PLAN_ENCODER, PLAN_GENERATOR, PLAN_SALIENCY, PLAN_RD, PLAN_ADAPTIVE = 1, 2, 3, 4, 5, 6
OK, ERR_INVALID, ERR_CUDA, ERR_MISSING, ERR_WORKSPACE = 0, -1, -2, -3, -4
SYM_MAX = 1023


class CicError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libcic error {code}: {message}")
        self.code = code


class cic_tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("h_data", C.POINTER(C.c_float)), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class cic_plan_opts(C.Structure):
    _fields_ = [("precision", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32), ("img_c", C.c_int32),
                ("latent_dim", C.c_int32), ("add_attention", C.c_int32), ("reserved", C.c_int32 * 8)]


class cic_adaptive_io(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_img", "d_mask", "d_bpp", "d_blended", "d_hq_latent_q", "d_lq_latent_q", "d_rd_params", "d_dt",
        "d_hq_symbols", "d_lq_symbols", "d_hq_latent", "d_lq_latent", "d_hq_scale", "d_lq_scale", "d_hq_out",
        "d_lq_out", "d_hq_ratio_sum")]


class cic_adaptive_state(C.Structure):
    _fields_ = [(n, C.c_void_p * 2) for n in ("x1", "x2", "x3", "x4_hi", "x4_lo", "g0")]


PHASE_ENCODE, PHASE_LATENT, PHASE_DECODE = 1, 2, 3

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  This package has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/cic.h declares
PROTOTYPES = {
    "cic_version": (_i, []),
    "cic_last_error": (C.c_char_p, []),
    "cic_device_info": (_i, [C.POINTER(_i)] * 3),
    "cic_plan_create": (_vp, [_i, C.POINTER(cic_tensor), _i, C.POINTER(cic_plan_opts)]),
    "cic_plan_destroy": (None, [_vp]),
    "cic_plan_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "cic_plan_last_launch_count": (_i, [_vp]),
    "cic_plan_set_profiling": (_i, [_vp, _i]),
    "cic_plan_get_profile": (_sz, [_vp, C.c_char_p, _sz]),
    "cic_autoencoder_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "cic_encoder_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "cic_generator_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "cic_saliency_forward": (_i, [_vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "cic_rd_forward": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "cic_adaptive_forward": (_i, [_vp, C.POINTER(cic_adaptive_io), _i, _i, _i, _vp, _sz, _vp]),
    "cic_adaptive_forward_phase": (_i, [_vp, C.POINTER(cic_adaptive_io), C.POINTER(cic_adaptive_state), _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "cic_conv2d_nhwc_f32": (_i, [_vp] * 6 + [_i] * 9 + [_vp]),
    "cic_conv2d_transpose4x4s2_nhwc_f32": (_i, [_vp] * 6 + [_i] * 6 + [_vp]),
    "cic_dense_workspace_bytes": (_sz, [_i, _i, _i]),
    "cic_dense_f32": (_i, [_vp] * 4 + [_i] * 4 + [_vp, _sz, _vp]),
    "cic_conv2d_tc_workspace_bytes": (_sz, [_i] * 10),
    "cic_conv2d_nhwc_tc": (_i, [_vp] * 7 + [_i] * 12 + [_vp, _sz, _vp]),
    "cic_dense_tc_workspace_bytes": (_sz, [_i, _i, _i]),
    "cic_dense_tc": (_i, [_vp] * 4 + [_i] * 5 + [_vp, _sz, _vp]),
    "cic_attention_workspace_bytes": (_sz, [_i, _i, _i]),
    "cic_self_attention_f32": (_i, [_vp] * 7 + [_f, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "cic_quantize_latent": (_i, [_vp] * 7 + [_i, _i, _vp]),
    "cic_rate_scalars": (_i, [_vp] * 4 + [_i, _vp]),
    "cic_roi_mask_blend": (_i, [_vp] * 7 + [_i, _i, _i, _vp]),
    "cic_hq_ratio_sweep": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "cic_symbol_entropy_bits": (_i, [_vp, _vp, _i, _i, _vp]),
    "cic_saliency_mask_workspace_bytes": (_sz, [_i, _i, _i]),
    "cic_saliency_mask_smooth": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "cic_saliency_enhance_workspace_bytes": (_sz, [_i, _i, _i]),
    "cic_saliency_enhance": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "cic_saliency_mask_binary_workspace_bytes": (_sz, [_i]),
    "cic_saliency_mask_binary": (_i, [_vp, _vp, C.c_double, _i, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "cic_saliency_map_workspace_bytes": (_sz, [_i, _i, _i]),
    "cic_saliency_map_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "cic_rans_max_bytes": (_sz, [_i, _i]),
    "cic_rans_workspace_bytes": (_sz, [_i, _i]),
    "cic_rans_encode": (_i, [_vp, _i, _i, _vp, _sz, _vp, _vp, _sz, _vp]),
    "cic_rans_decode": (_i, [_vp, _sz, _vp, _i, _i, _vp]),
    "cic_jpeg_max_bytes": (_sz, [_i, _i]),
    "cic_jpeg_workspace_bytes": (_sz, [_i, _i, _i]),
    "cic_jpeg_encode_u8": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _sz, _vp]),
    "cic_f32_to_u8_trunc": (_i, [_vp, _vp, _sz, _f, _vp]),
    "cic_u8_to_f32_signed": (_i, [_vp, _vp, _sz, _vp]),
    "cic_f32_signed_to_u8": (_i, [_vp, _vp, _sz, _vp]),
    "cic_metrics_psnr_ssim_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp]),
    "cic_metrics_psnr_ssim_f32_fast": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp]),
    "cic_msssim_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "cic_msssim_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _sz, _vp]),
    "cic_metric_sums": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "cic_metrics_psnr_ssim_gray_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.cic_last_error().decode("utf-8", "replace")


def check(code: int) -> None:
    if code != OK:
        raise CicError(code, last_error())
