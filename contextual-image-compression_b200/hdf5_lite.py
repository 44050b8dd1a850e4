"""A small pure-Python HDF5 reader - enough of the format to read Keras `.h5` checkpoints without h5py.

The reference saves its models with `model.save("*.h5")` (GAN_train.py:548-581) and reloads them with
`keras.models.load_model` (GAN_test.py:37-78, test_autoencoder.py:34).  h5py does not exist in this image, so this module reads
the container itself.  It implements the subset of the HDF5 File Format Specification (version 2.0 of the document, the "1.x"
on-disk structures) that libhdf5 writes with default settings - which is what h5py, and therefore Keras, produce:

  * superblock versions 0 and 1 (also behind a user block: the signature is searched at 0, 512, 1024, ...), 2 and 3;
  * old-style groups: symbol-table message -> B-tree v1 (node type 0) + local heap + symbol-table nodes (SNOD);
    new-style groups as far as compact link messages go (no dense link storage);
  * object headers version 1 (with continuation blocks) and version 2 ("OHDR" / "OCHK");
  * messages: dataspace (v1, v2), datatype (fixed point, floating point, fixed strings, variable-length strings through the global
    heap), data layout v3 (compact, contiguous, chunked with the deflate / shuffle filters through a v1 chunk B-tree) and the
    older layout v1 / v2, filter pipeline (v1, v2), attribute (v1, v2, v3), symbol table, link, continuation.

Pinned on a real file: scipy ships `testhdf5_7.4_GLNX86.mat`, an HDF5 file written by MATLAB's libhdf5 (user block, superblock v0,
symbol-table group, v1 object header, float64 dataset, string attribute); `tests/test_hdf5_lite.py` reads it and compares the
dataset with the same variable in the v5 MAT file next to it.  The Keras-specific layout on top (`keras_h5.py`) follows the Keras
source (legacy H5 format) and is exercised on files written by the minimal writer in the tests - no Keras-written file exists here.
"""
from __future__ import annotations

import mmap
import os
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


class _Reader:
    def __init__(self, data):
        self.d = data                        # bytes or a read-only mmap
        self.base = 0
        self.O = 8   # size of offsets
        self.L = 8   # size of lengths

    def u(self, pos: int, n: int) -> int:
        return int.from_bytes(self.d[pos:pos + n], "little")

    def off(self, pos: int) -> int:
        v = self.u(pos, self.O)
        return UNDEF if v == (1 << (8 * self.O)) - 1 else v

    def length(self, pos: int) -> int:
        return self.u(pos, self.L)

    def off_from(self, body: bytes, pos: int) -> int:
        """an offset-sized field of a message body (not of the file)"""
        return int.from_bytes(body[pos:pos + self.O], "little")

    def at(self, addr: int) -> int:
        """file position of an address stored in the file (addresses are relative to the base address)"""
        return self.base + addr


class Dataset:
    def __init__(self, f: "File", name: str, msgs: List[Tuple[int, int, bytes, int]]):
        self.file, self.name, self._msgs = f, name, msgs
        self.attrs = f._attributes(msgs)
        self.shape, self.dtype, self._kind = f._dataset_meta(msgs)

    def read(self) -> np.ndarray:
        return self.file._dataset_read(self._msgs, self.shape, self.dtype, self._kind)

    def __repr__(self):
        return f"<Dataset {self.name} {self.shape} {self.dtype}>"


class Group:
    def __init__(self, f: "File", name: str, header_addr: int):
        self.file, self.name = f, name
        self._msgs = f._object_header(header_addr)
        self.attrs = f._attributes(self._msgs)
        self._links = f._group_links(self._msgs)       # name -> object header address, in name order of the B-tree

    def keys(self) -> List[str]:
        return list(self._links)

    def __contains__(self, name: str) -> bool:
        return name in self._links

    def __getitem__(self, path: str):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(f"'{part}' not found in '{node.name}'")
            addr = node._links[part]
            msgs = self.file._object_header(addr)
            child = (node.name.rstrip("/") + "/" + part)
            if any(t == 0x0008 for t, _, _, _ in msgs):
                node = Dataset(self.file, child, msgs)
            else:
                node = Group(self.file, child, addr)
        return node

    def __repr__(self):
        return f"<Group {self.name} {self.keys()}>"


class File(Group):
    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray)):
            data = bytes(path_or_bytes)
        else:                                   # map the file: a full-size checkpoint is ~1 GB and only the tensors are copied out
            with open(path_or_bytes, "rb") as fh:
                size = os.fstat(fh.fileno()).st_size
                data = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) if size else b""
        self.r = _Reader(data)
        root = self._superblock()
        Group.__init__(self, self, "/", root)

    # ---- superblock -----------------------------------------------------------------------------------------------------------
    def _superblock(self) -> int:
        r, d = self.r, self.r.d
        pos = 0
        while pos + 8 <= len(d) and d[pos:pos + 8] != SIGNATURE:
            pos = 512 if pos == 0 else pos * 2
        if d[pos:pos + 8] != SIGNATURE:
            raise Hdf5Error("not an HDF5 file (no signature at 0, 512, 1024, ...)")
        ver = d[pos + 8]
        if ver in (0, 1):
            r.O, r.L = d[pos + 13], d[pos + 14]
            p = pos + 24 + (4 if ver == 1 else 0)
            base = r.u(p, r.O)
            r.base = base if base not in (0, UNDEF) or pos == 0 else pos
            if pos and base == 0:          # user block with a zero base address: addresses are relative to the superblock
                r.base = pos
            p += 4 * r.O                   # base, free-space, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            return r.off(p + r.O)
        if ver in (2, 3):
            r.O, r.L = d[pos + 9], d[pos + 10]
            base = r.u(pos + 12, r.O)
            r.base = base if base else pos
            return r.off(pos + 12 + 3 * r.O)
        raise Hdf5Error(f"superblock version {ver} is not supported")

    # ---- object headers -------------------------------------------------------------------------------------------------------
    def _object_header(self, addr: int) -> List[Tuple[int, int, bytes, int]]:
        """[(message type, flags, data, creation order)] of the object header at `addr`, continuation blocks included."""
        r, d = self.r, self.r.d
        pos = r.at(addr)
        msgs: List[Tuple[int, int, bytes, int]] = []
        if d[pos:pos + 4] == b"OHDR":
            return self._object_header_v2(pos)
        if d[pos] != 1:
            raise Hdf5Error(f"object header version {d[pos]} at {addr:#x} is not supported")
        nmsg = r.u(pos + 2, 2)
        size = r.u(pos + 8, 4)
        blocks = [(pos + 16, size)]        # 12 bytes of prefix, padded to 8
        while blocks and len(msgs) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, flags = r.u(p, 2), r.u(p + 2, 2), d[p + 4]
                body = d[p + 8:p + 8 + msize]
                if mtype == 0x0010:        # continuation
                    blocks.append((r.at(r.off_from(body, 0)), int.from_bytes(body[r.O:r.O + r.L], "little")))
                msgs.append((mtype, flags, body, len(msgs)))
                p += 8 + msize
        return msgs

    def _object_header_v2(self, pos: int) -> List[Tuple[int, int, bytes, int]]:
        r, d = self.r, self.r.d
        flags = d[pos + 5]
        p = pos + 6
        if flags & 0x20:
            p += 16                         # four time stamps
        if flags & 0x10:
            p += 4                          # attribute phase change values
        nsize = 1 << (flags & 3)
        chunk0 = r.u(p, nsize)
        p += nsize
        track = bool(flags & 0x04)
        msgs: List[Tuple[int, int, bytes, int]] = []
        blocks = [(p, chunk0)]
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 4 + (2 if track else 0) <= end:
                mtype, msize, mflags = d[p], r.u(p + 1, 2), d[p + 3]
                p += 4 + (2 if track else 0)
                body = d[p:p + msize]
                if mtype == 0x10:
                    cpos = r.at(int.from_bytes(body[:r.O], "little"))
                    clen = int.from_bytes(body[r.O:r.O + r.L], "little")
                    if d[cpos:cpos + 4] != b"OCHK":
                        raise Hdf5Error("bad object header continuation chunk")
                    blocks.append((cpos + 4, clen - 8))      # minus signature and checksum
                if mtype != 0:
                    msgs.append((mtype, mflags, body, len(msgs)))
                p += msize
        return msgs

    # ---- groups -----------------------------------------------------------------------------------------------------------------
    def _group_links(self, msgs) -> Dict[str, int]:
        r, d = self.r, self.r.d
        links: Dict[str, int] = {}
        for mtype, _, body, _ in msgs:
            if mtype == 0x0011:            # symbol table message: B-tree address, local heap address
                btree, heap = r.off_from(body, 0), r.off_from(body, r.O)
                hp = r.at(heap)
                if d[hp:hp + 4] != b"HEAP":
                    raise Hdf5Error("bad local heap")
                heap_data = r.at(r.off(hp + 8 + 2 * r.L))
                self._btree_group(btree, heap_data, links)
            elif mtype == 0x0006:          # link message (new-style group, compact storage)
                flags = body[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[p]; p += 1
                if flags & 0x04:
                    p += 8                 # creation order
                if flags & 0x10:
                    p += 1                 # charset
                nlen_size = 1 << (flags & 3)
                nlen = int.from_bytes(body[p:p + nlen_size], "little"); p += nlen_size
                name = body[p:p + nlen].decode("utf-8"); p += nlen
                if ltype == 0:
                    links[name] = int.from_bytes(body[p:p + r.O], "little")
            elif mtype == 0x0002 and len(body) >= 2:
                # link info: a fractal-heap address that is defined means dense link storage
                p = 2 + (8 if body[1] & 1 else 0)
                if int.from_bytes(body[p:p + r.O], "little") != (1 << (8 * r.O)) - 1:
                    raise Hdf5Error("dense link storage (fractal heap) is not supported: save the file with default h5py settings")
        return links

    def _btree_group(self, addr: int, heap_data: int, links: Dict[str, int]) -> None:
        r, d = self.r, self.r.d
        if addr == UNDEF:
            return
        pos = r.at(addr)
        if d[pos:pos + 4] != b"TREE" or d[pos + 4] != 0:
            raise Hdf5Error("bad group B-tree node")
        level, used = d[pos + 5], r.u(pos + 6, 2)
        p = pos + 8 + 2 * r.O
        for i in range(used):
            child = r.off(p + r.L)          # key i (heap offset), child i
            if level > 0:
                self._btree_group(child, heap_data, links)
            else:
                sp = r.at(child)
                if d[sp:sp + 4] != b"SNOD":
                    raise Hdf5Error("bad symbol table node")
                n = r.u(sp + 6, 2)
                ep = sp + 8
                for _ in range(n):
                    name_off, ohdr = r.off(ep), r.off(ep + r.O)
                    q = heap_data + name_off
                    name = d[q:d.find(b"\x00", q)].decode("utf-8")
                    links[name] = ohdr
                    ep += 2 * r.O + 24
            p += r.L + r.O

    # ---- datatypes, dataspaces -----------------------------------------------------------------------------------------------------
    def _datatype(self, body: bytes):
        """-> (kind, numpy dtype or None, element size).  kind: 'num', 'str' (fixed), 'vstr' (variable-length string)."""
        cls, bits0 = body[0] & 0x0F, body[1]
        size = int.from_bytes(body[4:8], "little")
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return "num", np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}"), size
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Error(f"{size}-byte floating point type is not supported")
            return "num", np.dtype(f"{order}f{size}"), size
        if cls == 3:
            return "str", np.dtype(f"S{size}"), size
        if cls == 9:
            vtype = bits0 & 0x0F
            if vtype == 1:
                return "vstr", None, size
            raise Hdf5Error("variable-length sequences are not supported")
        raise Hdf5Error(f"datatype class {cls} is not supported")

    def _dataspace(self, body: bytes) -> Tuple[int, ...]:
        ver, rank = body[0], body[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:               # null dataspace
                return (0,)
            p = 4
        else:
            raise Hdf5Error(f"dataspace version {ver} is not supported")
        L = self.r.L
        return tuple(int.from_bytes(body[p + i * L:p + (i + 1) * L], "little") for i in range(rank))

    def _decode(self, raw: bytes, shape, kind, dtype, esize) -> np.ndarray:
        n = int(np.prod(shape)) if shape else 1
        if kind == "vstr":
            r = self.r
            out = []
            step = 4 + r.O + 4
            for i in range(n):
                e = raw[i * step:(i + 1) * step]
                out.append(self._global_heap_object(int.from_bytes(e[4:4 + r.O], "little"), int.from_bytes(e[4 + r.O:], "little")))
            return np.array(out, dtype=object).reshape(shape)
        arr = np.frombuffer(raw[:n * esize], dtype=dtype).reshape(shape)
        return arr

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        r, d = self.r, self.r.d
        pos = r.at(addr)
        if d[pos:pos + 4] != b"GCOL":
            raise Hdf5Error("bad global heap collection")
        size = r.length(pos + 8)
        p, end = pos + 8 + r.L, pos + size
        while p + 8 + r.L <= end:
            idx, osize = r.u(p, 2), r.length(p + 8)
            if idx == index:
                return d[p + 8 + r.L:p + 8 + r.L + osize]
            if idx == 0:
                break
            p += 8 + r.L + ((osize + 7) & ~7)
        raise Hdf5Error("global heap object not found")

    # ---- attributes ----------------------------------------------------------------------------------------------------------------
    def _attributes(self, msgs) -> Dict[str, object]:
        out: Dict[str, object] = {}
        for mtype, _, body, _ in msgs:
            if mtype != 0x000C:
                continue
            ver = body[0]
            nsize, tsize, ssize = (int.from_bytes(body[2:4], "little"), int.from_bytes(body[4:6], "little"), int.from_bytes(body[6:8], "little"))
            p = 8 + (1 if ver == 3 else 0)
            pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
            name = body[p:p + nsize].split(b"\x00")[0].decode("utf-8"); p += pad(nsize)
            kind, dtype, esize = self._datatype(body[p:p + tsize]); p += pad(tsize)
            shape = self._dataspace(body[p:p + ssize]) if ssize else (); p += pad(ssize)
            val = self._decode(body[p:], shape, kind, dtype, esize)
            if val.shape == ():
                val = val[()]
            out[name] = val
        return out

    # ---- datasets ------------------------------------------------------------------------------------------------------------------
    def _dataset_meta(self, msgs):
        shape, kind, dtype, esize = (), None, None, 0
        for mtype, _, body, _ in msgs:
            if mtype == 0x0001:
                shape = self._dataspace(body)
            elif mtype == 0x0003:
                kind, dtype, esize = self._datatype(body)
        if kind is None:
            raise Hdf5Error("dataset without a datatype message")
        return shape, dtype, (kind, esize)

    def _dataset_read(self, msgs, shape, dtype, kind_esize) -> np.ndarray:
        r, d = self.r, self.r.d
        kind, esize = kind_esize
        layout = next((body for t, _, body, _ in msgs if t == 0x0008), None)
        if layout is None:
            raise Hdf5Error("dataset without a data layout message")
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * (esize if kind != "vstr" else 4 + r.O + 4)
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 0:                                   # compact
                size = int.from_bytes(layout[2:4], "little")
                raw = layout[4:4 + size]
            elif cls == 1:                                 # contiguous
                addr = int.from_bytes(layout[2:2 + r.O], "little")
                raw = b"\x00" * nbytes if addr == (1 << (8 * r.O)) - 1 else d[r.at(addr):r.at(addr) + nbytes]
            elif cls == 2:                                 # chunked
                rank = layout[2]
                btree = int.from_bytes(layout[3:3 + r.O], "little")
                cdims = tuple(int.from_bytes(layout[3 + r.O + 4 * i:7 + r.O + 4 * i], "little") for i in range(rank))
                return self._read_chunked(btree, shape, cdims[:-1], dtype, esize, msgs)
            else:
                raise Hdf5Error(f"data layout class {cls} is not supported")
        elif ver in (1, 2):
            rank, cls = layout[1], layout[2]
            p = 8
            addr = None
            if cls != 0:
                addr = int.from_bytes(layout[p:p + r.O], "little"); p += r.O
            dims = tuple(int.from_bytes(layout[p + 4 * i:p + 4 * i + 4], "little") for i in range(rank)); p += 4 * rank
            if cls == 0:
                size = int.from_bytes(layout[p:p + 4], "little")
                raw = layout[p + 4:p + 4 + size]
            elif cls == 1:
                raw = d[r.at(addr):r.at(addr) + nbytes]
            else:
                return self._read_chunked(addr, shape, dims[:-1], dtype, esize, msgs)
        else:
            raise Hdf5Error(f"data layout version {ver} is not supported")
        return self._decode(raw, shape, kind, dtype, esize)

    def _filters(self, msgs) -> List[int]:
        body = next((b for t, _, b, _ in msgs if t == 0x000B), None)
        if body is None:
            return []
        ver, nf = body[0], body[1]
        p = 8 if ver == 1 else 2
        ids = []
        for _ in range(nf):
            fid = int.from_bytes(body[p:p + 2], "little")
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(body[p + 2:p + 4], "little"); q = p + 8
            else:
                nlen = 0; q = p + 6
            ncd = int.from_bytes(body[q - 2:q], "little")
            q += (nlen + 7) & ~7 if ver == 1 else nlen
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            ids.append(fid)
            p = q
        return ids

    def _read_chunked(self, btree: int, shape, cdims, dtype, esize, msgs) -> np.ndarray:
        r, d = self.r, self.r.d
        filters = self._filters(msgs)
        for f in filters:
            if f not in (1, 2):
                raise Hdf5Error(f"filter {f} is not supported (only deflate and shuffle)")
        out = np.zeros(shape, dtype=dtype)
        rank = len(shape)

        def walk(addr):
            pos = r.at(addr)
            if d[pos:pos + 4] != b"TREE" or d[pos + 4] != 1:
                raise Hdf5Error("bad chunk B-tree node")
            level, used = d[pos + 5], r.u(pos + 6, 2)
            p = pos + 8 + 2 * r.O
            ksize = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize, fmask = r.u(p, 4), r.u(p + 4, 4)
                offs = tuple(r.u(p + 8 + 8 * i, 8) for i in range(rank))
                child = r.off(p + ksize)
                if level > 0:
                    walk(child)
                else:
                    raw = d[r.at(child):r.at(child) + csize]
                    for f in reversed(filters):
                        if f == 1 and not fmask & (1 << filters.index(f)):
                            raw = zlib.decompress(raw)
                        elif f == 2 and not fmask & (1 << filters.index(f)):
                            a = np.frombuffer(raw, np.uint8).reshape(esize, -1)
                            raw = a.T.tobytes()
                    chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cdims))).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
                p += ksize + r.O
        if btree != (1 << (8 * r.O)) - 1:
            walk(btree)
        return out

