"""Minimal reader for the reference's GMM model files (``*.jld``).

The reference stores models with JLD v0.1 (``bin/train_gmm.jl:106-113``): an HDF5 file with a
512-byte user block, superblock v0, old-style groups (TREE/HEAP/SNOD), v1 object headers and
un-filtered compact/contiguous datasets named ``weights``, ``means``, ``covars``, ``diff`` and
``n_components``.  Neither h5py nor libhdf5 exists in this image, so this walks exactly that
subset of the format with ``struct`` + ``numpy.frombuffer`` and refuses anything else.

Returned arrays have the Julia shapes -- ``means`` (2D, M), ``covars`` (2D, 2D, M),
``weights`` (M,) -- in column-major order, ready for :class:`GMMMap`.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class JLDFormatError(ValueError):
    pass


class _File:
    def __init__(self, buf: bytes):
        self.b = buf
        for base in (0, 512, 1024, 2048, 4096):
            if buf[base:base + 8] == _SIG:
                self.base = base
                break
        else:
            raise JLDFormatError("no HDF5 superblock found")
        sb = self.base
        if buf[sb + 8] != 0:
            raise JLDFormatError(f"superblock version {buf[sb + 8]} unsupported (want 0)")
        self.O, self.L = buf[sb + 13], buf[sb + 14]
        if (self.O, self.L) != (8, 8):
            raise JLDFormatError("only 8-byte offsets/lengths supported")
        # sig8 ver4 sizes4 K4 flags4 -> 24 ; base, freespace, eof, driver (4 x O) ; root entry
        self.addr_base = self.u64(sb + 24)
        root = sb + 24 + 4 * 8
        self.root_header = self.u64(root + 8)
        if self.u32(root + 16) != 1:
            raise JLDFormatError("root group without cached B-tree/heap addresses")
        self.root_btree, self.root_heap = self.u64(root + 24), self.u64(root + 32)

    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]
    def abs(self, a): return a + self.addr_base

    def heap_data(self, heap_addr):
        h = self.abs(heap_addr)
        if self.b[h:h + 4] != b"HEAP":
            raise JLDFormatError("bad local heap signature")
        return self.abs(self.u64(h + 24))

    def group_entries(self, btree_addr, heap_addr) -> Dict[str, int]:
        """name -> object header address for an old-style group."""
        data = self.heap_data(heap_addr)
        out: Dict[str, int] = {}

        def walk(addr):
            n = self.abs(addr)
            if self.b[n:n + 4] != b"TREE" or self.b[n + 4] != 0:
                raise JLDFormatError("bad group B-tree node")
            level, used = self.b[n + 5], self.u16(n + 6)
            p = n + 8 + 16  # skip siblings
            for i in range(used):
                child = self.u64(p + 8 + i * 16)  # key_i (8) then child_i (8)
                if level > 0:
                    walk(child)
                else:
                    s = self.abs(child)
                    if self.b[s:s + 4] != b"SNOD":
                        raise JLDFormatError("bad symbol node")
                    for k in range(self.u16(s + 6)):
                        e = s + 8 + k * 40
                        name_off, hdr = self.u64(e), self.u64(e + 8)
                        q = data + name_off
                        out[self.b[q:self.b.index(b"\0", q)].decode()] = hdr

        walk(btree_addr)
        return out

    def messages(self, hdr_addr):
        h = self.abs(hdr_addr)
        if self.b[h] != 1:
            raise JLDFormatError("only v1 object headers supported")
        nmsg, size = self.u16(h + 2), self.u32(h + 8)
        blocks = [(h + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, flags = self.u16(p), self.u16(p + 2), self.b[p + 4]
                body = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self.abs(self.u64(body)), self.u64(body + 8)))
                if flags & 0x02 and mtype == 0x03:
                    # shared message (JLD commits its datatypes under /_types): follow the
                    # reference to the committed datatype's own object header
                    ver = self.b[body]
                    ref = self.u64(body + (2 if ver >= 2 else 8))
                    tgt = [(b2, s2) for t2, b2, s2 in self.messages(ref) if t2 == 0x03]
                    msgs.append((mtype,) + (tgt[0] if tgt else (body, msize)))
                else:
                    msgs.append((mtype, body, msize))
                p = body + msize
        return msgs

    def dataset(self, hdr_addr) -> np.ndarray:
        dims, dtype, raw = None, None, None
        for mtype, p, msize in self.messages(hdr_addr):
            if mtype == 0x01:  # dataspace
                ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
                q = p + (8 if ver == 1 else 4)
                dims = [self.u64(q + 8 * i) for i in range(rank)]
            elif mtype == 0x03:  # datatype
                cls, size = self.b[p] & 0x0F, self.u32(p + 4)
                if self.b[p + 1] & 1:
                    raise JLDFormatError("big-endian data unsupported")
                if cls == 1 and size == 8:
                    dtype = np.dtype("<f8")
                elif cls == 0:
                    signed = bool(self.b[p + 1] & 0x08)
                    dtype = np.dtype(f"<{'i' if signed else 'u'}{size}")
                elif cls in (4, 5) and size == 1:
                    dtype = np.dtype("u1")  # JLD writes Bool as a committed 1-byte bitfield/opaque
                else:
                    dtype = None  # strings / compound (JLD bookkeeping) -- not needed
            elif mtype == 0x08:  # layout
                if self.b[p] != 3:
                    raise JLDFormatError(f"data layout version {self.b[p]} unsupported")
                lclass = self.b[p + 1]
                if lclass == 0:
                    n = self.u16(p + 2)
                    raw = (p + 4, n)
                elif lclass == 1:
                    raw = (self.abs(self.u64(p + 2)), self.u64(p + 10))
                else:
                    raise JLDFormatError("chunked/filtered datasets unsupported")
        if dtype is None or raw is None or dims is None:
            raise JLDFormatError("not a plain numeric dataset")
        count = int(np.prod(dims)) if dims else 1
        arr = np.frombuffer(self.b, dtype=dtype, count=count, offset=raw[0])
        # HDF5 dims are row-major; Julia wrote its column-major array with reversed dims.
        return arr.reshape(dims[::-1], order="F") if dims else arr.reshape(())


def load(path: str) -> Dict[str, object]:
    """``JLD.load(path)`` for the reference's model schema.

    Returns ``{"weights": (M,), "means": (2D, M), "covars": (2D, 2D, M), "diff": bool,
    "n_components": int}``; datasets that are not plain numerics (JLD's ``_creator`` group
    etc.) are skipped.
    """
    with open(path, "rb") as f:
        hf = _File(f.read())
    out: Dict[str, object] = {}
    for name, hdr in hf.group_entries(hf.root_btree, hf.root_heap).items():
        if name.startswith("_"):
            continue
        try:
            a = hf.dataset(hdr)
        except JLDFormatError:
            continue
        if name == "diff":
            out[name] = bool(a.reshape(-1)[0])
        elif name == "n_components":
            out[name] = int(a.reshape(-1)[0])
        else:
            out[name] = np.asfortranarray(a.astype(np.float64))
    for k in ("weights", "means", "covars"):
        if k not in out:
            raise JLDFormatError(f"{path}: dataset '{k}' not found")
    return out
