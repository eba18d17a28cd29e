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


# ---- writer ---------------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _datatype_msg(dt: np.dtype) -> bytes:
    if dt == np.dtype("<f8"):      # IEEE binary64, little endian: class 1 (floating point), version 1
        return _msg(0x03, struct.pack("<B3BI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023))
    if dt == np.dtype("<i8"):      # class 0 (fixed point), signed
        return _msg(0x03, struct.pack("<B3BI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64))
    if dt == np.dtype("u1"):       # class 0, unsigned byte (the reader also accepts JLD's committed Bool type)
        return _msg(0x03, struct.pack("<B3BI", 0x10, 0x00, 0, 0, 1) + struct.pack("<HH", 0, 8))
    raise JLDFormatError(f"dtype {dt} not supported by the writer")


def save(path: str, weights, means, covars, diff: bool = False, n_components: int = None) -> None:
    """Writes a model in the schema of ``bin/train_gmm.jl:106-113`` (keys ``weights``, ``means``,
    ``covars``, ``diff``, ``n_components``) as the HDF5 subset JLD v0.1 uses for plain arrays: 512-byte
    user block with the JLD banner, superblock v0, one old-style root group (B-tree + local heap +
    symbol node), v1 object headers, contiguous little-endian datasets whose HDF5 dimensions are the
    reversed Julia dimensions (Julia writes its column-major memory as is).

    ``load`` reads the result back bit for bit (tests/test_host_logic.py round-trips both models of
    the reference's test/models).  The committed ``Bool`` datatype JLD attaches to ``diff`` and its
    ``_creator`` group are bookkeeping of the Julia package and are not reproduced: ``diff`` is stored
    as an unsigned byte.  Neither libhdf5 nor Julia exists in the build image, so reading these files
    with JLD.jl itself is untested."""
    w = np.ascontiguousarray(np.asarray(weights, dtype="<f8").reshape(-1))
    mu = np.asfortranarray(np.asarray(means, dtype="<f8"))
    sg = np.asfortranarray(np.asarray(covars, dtype="<f8"))
    if mu.ndim != 2 or sg.ndim != 3 or sg.shape != (mu.shape[0], mu.shape[0], mu.shape[1]) or w.shape != (mu.shape[1],):
        raise JLDFormatError("expected weights (M,), means (2D, M), covars (2D, 2D, M)")
    M = mu.shape[1]
    items = {      # name -> (HDF5 dims = reversed Julia dims, dtype, raw bytes in Julia memory order)
        "covars": (sg.shape[::-1], np.dtype("<f8"), sg.tobytes(order="F")),
        "diff": ((), np.dtype("u1"), bytes([1 if diff else 0])),
        "means": (mu.shape[::-1], np.dtype("<f8"), mu.tobytes(order="F")),
        "n_components": ((), np.dtype("<i8"), struct.pack("<q", M if n_components is None else int(n_components))),
        "weights": (w.shape, np.dtype("<f8"), w.tobytes()),
    }
    names = sorted(items)                                   # symbol-node entries are ordered by name
    LEAF_K, INTERNAL_K = 4, 16
    if len(names) > 2 * LEAF_K:
        raise JLDFormatError("too many datasets for one symbol node")
    # ---- local heap data segment: offset 0 holds the empty string, names are 8-byte aligned
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for nme in names:
        name_off[nme] = len(heap)
        heap += _pad8(nme.encode() + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 32) + b"\0" * 16          # one free block (next = 1: end of list)
    # ---- addresses (relative to the base address = size of the user block)
    SUPER, ROOT_HDR = 0, 96
    BTREE = ROOT_HDR + 16 + 24
    btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    HEAP = BTREE + btree_size
    HEAP_DATA = HEAP + 32
    SNOD = HEAP_DATA + len(heap)
    pos = SNOD + 8 + 2 * LEAF_K * 40
    hdr_addr, hdr_bytes, data_addr = {}, {}, {}
    for nme in names:                                       # object headers first, raw data after them
        dims, dt, raw = items[nme]
        hdr_addr[nme] = pos
        space = _msg(0x01, struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims))
        fill = _msg(0x05, struct.pack("<BBBB", 2, 2, 0, 0))
        layout_len = len(_msg(0x08, struct.pack("<BBQQ", 3, 1, 0, 0)))
        hdr_bytes[nme] = (space, _datatype_msg(dt), fill, layout_len)
        pos += 16 + len(space) + len(_datatype_msg(dt)) + len(fill) + layout_len
    for nme in names:
        pos += -pos % 8
        data_addr[nme] = pos
        pos += len(items[nme][2])
    eof = pos
    # ---- assemble
    BASE = 512
    out = bytearray(BASE + eof)
    banner = b"Julia data file (HDF5), version 0.1.0"
    out[:len(banner)] = banner

    def put(addr, b):
        out[BASE + addr:BASE + addr + len(b)] = b

    undef = 0xFFFFFFFFFFFFFFFF
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", BASE, undef, BASE + eof, undef)
    sb += struct.pack("<QQII", 0, ROOT_HDR, 1, 0) + struct.pack("<QQ", BTREE, HEAP)     # root entry caches B-tree / heap
    put(SUPER, sb)
    symtab = _msg(0x11, struct.pack("<QQ", BTREE, HEAP))
    put(ROOT_HDR, struct.pack("<BBHII4x", 1, 0, 1, 1, len(symtab)) + symtab)
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, undef, undef)
    bt += struct.pack("<Q", 0) + struct.pack("<Q", SNOD) + struct.pack("<Q", name_off[names[-1]])
    put(BTREE, bt)
    put(HEAP, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, HEAP_DATA))
    put(HEAP_DATA, bytes(heap))
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for nme in names:
        sn += struct.pack("<QQII16x", name_off[nme], hdr_addr[nme], 0, 0)
    put(SNOD, sn)
    for nme in names:
        space, dtm, fill, _ = hdr_bytes[nme]
        layout = _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr[nme], len(items[nme][2])))
        body = space + dtm + fill + layout
        put(hdr_addr[nme], struct.pack("<BBHII4x", 1, 0, 4, 1, len(body)) + body)
        put(data_addr[nme], items[nme][2])
    with open(path, "wb") as f:
        f.write(bytes(out))
