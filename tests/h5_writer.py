"""Test infrastructure: a minimal HDF5 WRITER producing the on-disk structures libhdf5 writes by default (superblock v0, symbol-table
groups = B-tree v1 + local heap + SNOD, version-1 object headers with an optional continuation block, version-1 attributes,
contiguous datasets), and on top of it the Keras legacy-H5 layout (`write_keras`).  It exists because neither h5py nor Keras is
available here; `hdf5_lite` (the product's reader) is pinned separately on a real libhdf5-written file, this writer only lets the
Keras-layout logic of `keras_h5` be exercised end to end through actual files.  Follows the HDF5 File Format Specification 2.0.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dataspace(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        exp_bits, man_bits, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        return struct.pack("<BBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize) + struct.pack(
            "<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)            # null-padded ASCII
    raise TypeError(dt)


def _message(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


class VlenStr(str):
    """an attribute value to be written as a variable-length UTF-8 string (what h5py does for a Python str): global heap + v3 attribute"""


def _attribute(name: str, value, writer=None) -> bytes:
    if isinstance(value, VlenStr):
        raw = value.encode("utf-8")
        # global heap collection with one object (index 1) and the free-space object (index 0)
        obj = struct.pack("<HH4xQ", 1, 1, len(raw)) + _pad8(raw)
        size = 16 + len(obj) + 16
        gaddr = writer.alloc(b"GCOL" + struct.pack("<B3xQ", 1, size) + obj + struct.pack("<HH4xQ", 0, 0, 0))
        nm = name.encode() + b"\x00"
        dt = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)
        ds = struct.pack("<BBBB", 2, 0, 0, 0)                                  # version 2, rank 0, scalar
        data = struct.pack("<IQI", len(raw), gaddr, 1)
        return _message(0x000C, struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(ds), 1) + nm + dt + ds + data)
    arr = np.asarray(value)
    if arr.dtype.kind == "U":
        arr = np.char.encode(arr, "utf-8")
    nm = name.encode() + b"\x00"
    dt, ds = _datatype(arr.dtype), _dataspace(arr.shape)
    return _message(0x000C, struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes())


class Chunked:
    def __init__(self, arr, chunks, deflate=True, shuffle=False):
        self.arr, self.chunks, self.deflate, self.shuffle = np.asarray(arr), tuple(chunks), deflate, shuffle


class Writer:
    def __init__(self, user_block: int = 0):
        self.user_block = user_block
        self.buf = bytearray(96)                 # superblock v0 with 8-byte offsets and lengths

    def alloc(self, data: bytes) -> int:
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _header(self, messages: List[bytes], split: bool) -> int:
        """version-1 object header; with `split`, the second half of the messages lives in a continuation block"""
        if split and len(messages) > 1:
            half = len(messages) // 2
            cont = b"".join(messages[half:])
            caddr = self.alloc(cont)
            first = messages[:half] + [_message(0x0010, struct.pack("<QQ", caddr, len(cont)))]
            n = len(messages) + 1
        else:
            first, n = messages, len(messages)
        body = b"".join(first)
        return self.alloc(struct.pack("<BxHII4x", 1, n, 1, len(body)) + body)

    def dataset(self, arr: np.ndarray, attrs: Dict[str, object] = None, split: bool = False) -> int:
        arr = np.ascontiguousarray(arr)
        daddr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        msgs = [_message(0x0001, _dataspace(arr.shape)), _message(0x0003, _datatype(arr.dtype)),
                _message(0x0008, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes))]
        msgs += [_attribute(k, v, self) for k, v in (attrs or {}).items()]
        return self._header(msgs, split)

    def chunked_dataset(self, arr: np.ndarray, chunks, deflate: bool = True, shuffle: bool = False) -> int:
        """chunked layout (v3 class 2) behind a one-node v1 chunk B-tree; filter pipeline v1: [shuffle,] deflate"""
        import itertools
        import zlib
        arr = np.ascontiguousarray(arr)
        rank, es = arr.ndim, arr.dtype.itemsize
        entries = []
        for idx in itertools.product(*[range(0, s, c) for s, c in zip(arr.shape, chunks)]):
            chunk = np.zeros(chunks, arr.dtype)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(idx, chunks, arr.shape))
            chunk[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
            raw = chunk.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()
            if deflate:
                raw = zlib.compress(raw, 4)
            entries.append((idx, self.alloc(raw), len(raw)))
        assert len(entries) <= 64
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
        for idx, addr, size in entries:
            node += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in idx) + struct.pack("<Q", 0) + struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)      # final key
        baddr = self.alloc(node)
        layout = struct.pack("<BBBQ", 3, 2, rank + 1, baddr) + b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", es)
        filt = []
        if shuffle:
            filt.append(struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I", es) + b"\x00" * 4)
        if deflate:
            filt.append(struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", 4) + b"\x00" * 4)
        msgs = [_message(0x0001, _dataspace(arr.shape)), _message(0x0003, _datatype(arr.dtype)), _message(0x0008, layout)]
        if filt:
            msgs.append(_message(0x000B, struct.pack("<BB6x", 1, len(filt)) + b"".join(filt)))
        return self._header(msgs, False)

    def group(self, children: Dict[str, int], attrs: Dict[str, object] = None, split: bool = False) -> int:
        names = sorted(children, key=lambda s: s.encode())
        heap = bytearray(8)                      # offset 0: the empty string
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode() + b"\x00")
        heap += b"\x00" * 16
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), len(heap) - 16, heap_data))
        snods = []
        for i in range(0, len(names), 8):        # group leaf node K = 4: at most 8 symbols per node
            part = names[i:i + 8]
            node = b"SNOD" + struct.pack("<BxH", 1, len(part))
            for n in part:
                node += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
            node += b"\x00" * (40 * (8 - len(part)))
            snods.append((self.alloc(node), offs[part[-1]]))
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for addr, last in snods:
            tree += struct.pack("<QQ", addr, last)
        tree += b"\x00" * (16 * (32 - len(snods)))
        tree_addr = self.alloc(tree)
        msgs = [_message(0x0011, struct.pack("<QQ", tree_addr, heap_addr))] + [_attribute(k, v, self) for k, v in (attrs or {}).items()]
        return self._header(msgs, split)

    def finish(self, root: int) -> bytes:
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, 4, 16, 0)
        sb += struct.pack("<QQQQ", self.user_block, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII16x", 0, root, 0, 0)
        assert len(sb) == 96, len(sb)
        self.buf[:96] = sb
        return b"\x00" * self.user_block + bytes(self.buf)


def _tree(w: Writer, node, attrs=None, split=False) -> int:
    """node: ndarray (dataset) or {name: node | (node, attrs)}"""
    if isinstance(node, Chunked):
        return w.chunked_dataset(node.arr, node.chunks, node.deflate, node.shuffle)
    if isinstance(node, np.ndarray):
        return w.dataset(node, attrs, split)
    children = {}
    for k, v in node.items():
        sub, a = v if isinstance(v, tuple) else (v, None)
        children[k] = _tree(w, sub, a, split)
    return w.group(children, attrs, split)


def write_tree(tree: dict, attrs=None, user_block: int = 0, split: bool = False) -> bytes:
    w = Writer(user_block)
    return w.finish(_tree(w, tree, attrs, split))


def write_keras(layers: Sequence[Tuple[str, str, Sequence[Tuple[str, np.ndarray]]]], model_name="model", with_config=True,
                split=False, user_block=0, weights_only=False) -> bytes:
    """layers: (class_name, layer_name, [(weight name as Keras writes it, e.g. 'conv2d/kernel:0', array)])"""
    mw = {}
    for cls, lname, ws in layers:
        g: dict = {}
        for wname, arr in ws:
            node = g
            parts = wname.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = np.asarray(arr)
        width = max([len(n.encode()) for n, _ in ws] + [1])
        wn = np.array([n.encode() for n, _ in ws], dtype=f"S{width}") if ws else np.zeros((0,), "S1")
        mw[lname] = (g, {"weight_names": wn})
    width = max(len(l[1].encode()) for l in layers)
    mw_attrs = {"layer_names": np.array([l[1].encode() for l in layers], dtype=f"S{width}"),
                "backend": np.bytes_(b"tensorflow"), "keras_version": np.bytes_(b"2.11.0")}
    if weights_only:
        return write_tree(mw, mw_attrs, user_block, split)
    root_attrs = {"keras_version": np.bytes_(b"2.11.0"), "backend": np.bytes_(b"tensorflow")}
    if with_config:
        cfg = {"class_name": "Functional", "config": {"name": model_name, "layers": [
            {"class_name": cls, "name": lname, "config": {"name": lname}, "inbound_nodes": []} for cls, lname, _ in layers]}}
        root_attrs["model_config"] = np.bytes_(json.dumps(cfg).encode())
    return write_tree({"model_weights": (mw, mw_attrs)}, root_attrs, user_block, split)
