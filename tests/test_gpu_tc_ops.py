"""GPU parity of the tcgen05 (tensor-core) operators against the CPU oracle in float64, through the C ABI.

Two arithmetic modes (include/cic.h):
  split=False  single-pass bf16 operands, fp32 accumulate: compared with the oracle evaluated on the
               bf16-rounded operands, so only the accumulation order differs (tight tolerance);
  split=True   3-term split-bf16: compared with the oracle on the original fp32 operands; the dropped
               lo*lo term bounds the relative error of every product by ~2^-16.
"""
import numpy as np
import pytest
import torch

from oracle import graphs

pytestmark = pytest.mark.gpu

ACTS = {"relu": torch.relu, "lrelu": graphs.lrelu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, None: lambda v: v}


def bf16_round(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def tol(split: bool, K: int):
    # split: |err| <~ 2^-16 * sum|x||w| ~ 1.5e-5 * sqrt(K) * rms; single pass on rounded operands: fp32 accumulation only
    return (4e-5 * max(1.0, np.sqrt(K / 256.0))) if split else 2e-4


CONV_CASES = [  # kh, stride, B, H, W, Cin, Cin2, Cout, act
    (3, 1, 2, 16, 24, 64, 0, 64, "relu"),        # autoencoder conv3 / conv_x2
    (3, 1, 1, 12, 20, 64, 64, 32, "relu"),       # autoencoder conv5: concat of two sources
    (3, 1, 2, 16, 16, 32, 0, 64, "relu"),        # 32-channel source: 64-byte swizzle K blocks
    (3, 1, 3, 9, 7, 32, 0, 32, "relu"),          # odd sizes: TMA zero fill on every edge
    (3, 1, 2, 16, 16, 32, 32, 3, "sigmoid"),     # autoencoder conv_out: N = 3 padded to 16
    (4, 1, 2, 16, 16, 32, 0, 3, "tanh"),         # generator conv_out: k4 'same' pads 1 / 2
    (4, 2, 2, 32, 32, 64, 0, 128, "lrelu"),      # encoder conv2
    (4, 2, 3, 16, 16, 128, 0, 256, "lrelu"),     # encoder conv3
    (4, 2, 5, 8, 8, 256, 0, 512, "lrelu"),       # encoder conv4: tile spans batch items
    (3, 2, 2, 16, 16, 32, 0, 64, "lrelu"),       # RD conv2: k3 s2 pads 0 / 1
    (1, 1, 2, 8, 8, 256, 0, 64, None),           # attention q|k projection
    (1, 1, 1, 32, 32, 256, 0, 256, None),
    (3, 1, 1, 40, 136, 64, 0, 64, "relu"),       # W > 128: several tiles per row, ragged last tile
    (3, 1, 1, 17, 128, 64, 0, 128, "relu"),      # CTA-pair kernel: 17 M tiles -> odd pair count, phantom tile masked
    (3, 1, 2, 32, 64, 128, 0, 256, "lrelu"),     # CTA-pair kernel: N tile 256, 32 M tiles
]


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("kh,stride,B,H,W,Cin,Cin2,Cout,act", CONV_CASES)
def test_conv2d_tc_matches_oracle(cic, split, kh, stride, B, H, W, Cin, Cin2, Cout, act):
    rng = np.random.default_rng(kh * 1000 + Cin * 7 + Cout + H)
    ct = Cin + Cin2
    x = rng.standard_normal((B, H, W, ct)).astype(np.float32)
    k = (rng.standard_normal((kh, kh, ct, Cout)) / np.sqrt(kh * kh * ct)).astype(np.float32)
    b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
    scale = (0.5 + rng.random(Cout)).astype(np.float32)
    shift = (rng.random(Cout) - 0.5).astype(np.float32)
    x1 = np.ascontiguousarray(x[..., :Cin])
    x2 = np.ascontiguousarray(x[..., Cin:]) if Cin2 else None
    got = cic.ops.conv2d_tc(x1, k, b, stride=stride, activation=act, scale=scale, shift=shift, x2=x2, split=split).cpu().numpy()
    xr, kr = (x, k) if split else (bf16_round(x), bf16_round(k))
    t = graphs.conv2d_same(graphs._nchw(torch.from_numpy(xr).double()), kr.astype(np.float64), b.astype(np.float64), stride, torch.float64)
    t = t * torch.from_numpy(scale).double().view(1, -1, 1, 1) + torch.from_numpy(shift).double().view(1, -1, 1, 1)
    want = graphs._nhwc(ACTS[act](t)).numpy()
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err < tol(split, kh * kh * ct), f"max-abs {err}"


DECONV_CASES = [(2, 4, 4, 512, 0, 256), (1, 8, 8, 256, 256, 128), (3, 8, 6, 64, 64, 32), (2, 16, 16, 128, 128, 64),
                (2, 16, 16, 128, 128, 128),   # CTA-pair kernel with four output phases
                (1, 34, 64, 64, 0, 256),      # CTA-pair kernel, 17 M tiles per phase
                (1, 34, 64, 64, 0, 64),       # merged-phase pair kernel (Cout 64): 17 M tiles -> phantom tile masked
                (4, 32, 32, 128, 128, 64),    # merged-phase pair kernel: generator deconv3 shape, two sources, 32 M tiles
                (2, 32, 64, 128, 0, 32)]      # merged-phase pair kernel, Cout 32 (generator deconv4 shape)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("B,H,W,Cin,Cin2,Cout", DECONV_CASES)
def test_conv_transpose_tc_matches_oracle(cic, split, B, H, W, Cin, Cin2, Cout):
    rng = np.random.default_rng(H * Cin + Cout)
    ct = Cin + Cin2
    x = rng.standard_normal((B, H, W, ct)).astype(np.float32)
    k = (rng.standard_normal((4, 4, Cout, ct)) / np.sqrt(4 * ct)).astype(np.float32)
    b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
    x1 = np.ascontiguousarray(x[..., :Cin])
    x2 = np.ascontiguousarray(x[..., Cin:]) if Cin2 else None
    got = cic.ops.conv2d_tc(x1, k, b, activation="lrelu", x2=x2, transpose=True, split=split).cpu().numpy()
    xr, kr = (x, k) if split else (bf16_round(x), bf16_round(k))
    t = graphs.lrelu(graphs.conv2d_transpose_same_k4s2(graphs._nchw(torch.from_numpy(xr).double()), kr.astype(np.float64),
                                                       b.astype(np.float64), torch.float64))
    want = graphs._nhwc(t).numpy()
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err < tol(split, 4 * ct), f"max-abs {err}"


_DC2_SCRIPT = r"""
import sys
import numpy as np, torch
sys.path.insert(0, {root!r})
import cic_b200 as cic
from oracle import graphs
bf = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()
for (B, H, W, Cin, Cin2, Cout) in [(4, 32, 32, 128, 128, 64), (1, 34, 64, 64, 0, 64), (2, 32, 64, 128, 0, 32)]:
    rng = np.random.default_rng(H * Cin + Cout)
    ct = Cin + Cin2
    x = rng.standard_normal((B, H, W, ct)).astype(np.float32)
    k = (rng.standard_normal((4, 4, Cout, ct)) / np.sqrt(4 * ct)).astype(np.float32)
    b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
    x1 = np.ascontiguousarray(x[..., :Cin]); x2 = np.ascontiguousarray(x[..., Cin:]) if Cin2 else None
    got = cic.ops.conv2d_tc(x1, k, b, activation="lrelu", x2=x2, transpose=True, split=False).cpu().numpy()
    t = graphs.lrelu(graphs.conv2d_transpose_same_k4s2(graphs._nchw(torch.from_numpy(bf(x)).double()), bf(k).astype(np.float64),
                                                       b.astype(np.float64), torch.float64))
    err = np.abs(got - graphs._nhwc(t).numpy()).max()
    assert err < 2e-4, (B, H, W, Cin, Cin2, Cout, err)
print("DC2_OK")
"""


def test_merged_phase_pair_deconv_kernel(cic):
    """tc_deconv2_kernel (tc_gemm2.cu) is off by default; CIC_TC_DC2 is read once per process, so it is exercised in a child
    process: the generator's deconv3 / deconv4 shapes and an odd tile count (phantom tile) against the float64 oracle."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CIC_TC_DC2="2")
    r = subprocess.run([sys.executable, "-c", _DC2_SCRIPT.format(root=root)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DC2_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("B,K,N", [(1, 8192, 64), (5, 131072, 32), (256, 1024, 2048), (3, 64, 8192), (4, 32, 1024), (130, 512, 3)])
def test_dense_tc_matches_oracle(cic, split, B, K, N):
    rng = np.random.default_rng(K + N)
    x = rng.standard_normal((B, K)).astype(np.float32)
    k = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    got = cic.ops.dense_tc(x, k, b, activation="relu", split=split).cpu().numpy()
    xr, kr = (x, k) if split else (bf16_round(x), bf16_round(k))
    want = np.maximum(xr.astype(np.float64) @ kr.astype(np.float64) + b, 0)
    err = np.abs(got - want).max()
    assert err < tol(split, K) * (3 if K > 65536 else 1), f"max-abs {err}"
    again = cic.ops.dense_tc(x, k, b, activation="relu", split=split).cpu().numpy()
    np.testing.assert_array_equal(got, again)                                   # fixed-order split-K: bit-reproducible


def test_conv2d_tc_rejects_bad_channels(cic):
    x = np.zeros((1, 8, 8, 3), np.float32)
    k = np.zeros((3, 3, 3, 32), np.float32)
    with pytest.raises(cic.CicError):
        cic.ops.conv2d_tc(x, k)
