"""CPU restatement of the latent-symbol entropy coder (contextual-image-compression_b200/csrc/rans.cu).  TEST INFRASTRUCTURE ONLY.

The reference has no entropy coder (its bitrate is nominal: GAN_test.py:310-325), so there is nothing of the reference's to
restate here; this file pins the GPU coder's byte stream: the same integer model rule and the same interleaved rANS, written
with numpy (32 lanes per row as vectors), must produce identical bytes, and `decode` must invert both.
Format: see include/cic.h (cic_rans_encode).
"""
from __future__ import annotations

import numpy as np

PROB_BITS = 14
M = 1 << PROB_BITS
L_LOW = 1 << 16
SYM_MAX = 1023
ALPHA = 2 * SYM_MAX + 1
HEADER, TABLE_BYTES = 32, 4096
MAGIC = 0x52434943


def build_freq(symbols: np.ndarray) -> np.ndarray:
    """counts -> frequencies summing to 2^14: f = max(1, floor(count * (M - K) / total)) for the K present symbols, the remainder
    goes to the most frequent symbol (lowest index on ties)."""
    s = np.clip(np.asarray(symbols, np.int64), -SYM_MAX, SYM_MAX) + SYM_MAX
    cnt = np.bincount(s.ravel(), minlength=ALPHA).astype(np.int64)
    total = int(cnt.sum())
    f = np.zeros(ALPHA, np.int64)
    if total == 0:
        f[SYM_MAX] = M
        return f
    k = int((cnt > 0).sum())
    present = cnt > 0
    f[present] = np.maximum(1, cnt[present] * (M - k) // total)
    best = int(np.argmax(cnt))              # first maximum = lowest index
    f[best] += M - int(f.sum())
    return f


def encode(symbols: np.ndarray) -> bytes:
    sym = np.asarray(symbols, np.int64)
    rows, L = sym.shape
    f = build_freq(sym)
    cum = np.concatenate([[0], np.cumsum(f)[:-1]])
    s = np.clip(sym, -SYM_MAX, SYM_MAX) + SYM_MAX
    T = (L + 31) // 32
    lanes = np.arange(32)
    row_blobs = []
    for r in range(rows):
        x = np.full(32, L_LOW, np.uint64)
        out_rev = []                                     # words in the order written (backwards in memory)
        for t in range(T - 1, -1, -1):
            idx = t * 32 + lanes
            active = idx < L
            sy = np.where(active, s[r, np.minimum(idx, L - 1)], 0)
            fr = np.where(active, f[sy], 1).astype(np.uint64)
            cu = np.where(active, cum[sy], 0).astype(np.uint64)
            emit = active & (x >= (fr << np.uint64(32 - PROB_BITS)))
            words = (x[emit] & np.uint64(0xFFFF)).astype(np.uint16)       # increasing lane order = increasing address
            out_rev.append(words)
            x = np.where(emit, x >> np.uint64(16), x)
            x = np.where(active, ((x // fr) << np.uint64(PROB_BITS)) + (x % fr) + cu, x)
        words = np.concatenate(out_rev[::-1]) if out_rev else np.zeros(0, np.uint16)   # later steps sit at lower addresses
        blob = x.astype("<u4").tobytes() + words.astype("<u2").tobytes()
        if len(words) & 1:
            blob += b"\x00\x00"
        row_blobs.append(blob)
    offsets = np.zeros(rows + 1, "<u4")
    offsets[1:] = np.cumsum([len(b) for b in row_blobs])
    header = np.array([MAGIC, 1, rows, L, PROB_BITS, ALPHA, 0, 0], "<u4").tobytes()
    table = np.zeros(2048, "<u2")
    table[:ALPHA] = f
    return header + table.tobytes() + offsets.tobytes() + b"".join(row_blobs)


def decode(stream: bytes) -> np.ndarray:
    buf = np.frombuffer(stream, np.uint8)
    hd = np.frombuffer(stream[:HEADER], "<u4")
    assert hd[0] == MAGIC and hd[1] == 1 and hd[4] == PROB_BITS and hd[5] == ALPHA, "not a CICR v1 stream"
    rows, L = int(hd[2]), int(hd[3])
    f = np.frombuffer(stream[HEADER:HEADER + TABLE_BYTES], "<u2")[:ALPHA].astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(f)[:-1]])
    lut = np.repeat(np.arange(ALPHA), f)
    assert len(lut) == M
    off0 = HEADER + TABLE_BYTES
    offsets = np.frombuffer(stream[off0:off0 + 4 * (rows + 1)], "<u4").astype(np.int64)
    payload = off0 + 4 * (rows + 1)
    out = np.zeros((rows, L), np.int32)
    T = (L + 31) // 32
    lanes = np.arange(32)
    for r in range(rows):
        a, b = payload + offsets[r], payload + offsets[r + 1]
        x = np.frombuffer(buf[a:a + 128].tobytes(), "<u4").astype(np.uint64)
        words = np.frombuffer(buf[a + 128:b].tobytes(), "<u2").astype(np.uint64)
        rp = 0
        for t in range(T):
            idx = t * 32 + lanes
            active = idx < L
            slot = (x & np.uint64(M - 1)).astype(np.int64)
            sy = lut[slot]
            xn = f[sy].astype(np.uint64) * (x >> np.uint64(PROB_BITS)) + slot.astype(np.uint64) - cum[sy].astype(np.uint64)
            x = np.where(active, xn, x)
            out[r, idx[active]] = sy[active] - SYM_MAX
            need = active & (x < np.uint64(L_LOW))
            n = int(need.sum())
            if n:
                x[need] = (x[need] << np.uint64(16)) | words[rp:rp + n]
                rp += n
    return out
