"""Minimal read-only HDF5 access for the dataset files `gen.py` writes (gen.py:406-413), used when `h5py` is not importable.

The reference reads its training / inference data with h5py (dataloader.py:73-105, inference.py:59-122): variable-length
float32 / float64 rows (`signal`, `noise`, `noisy`, `times`), fixed 1-D arrays (`lengths`, `t_merger`, `mass1`, ...), optional
per-sample PSD rows, and scalar root attributes (`sampling_rate`, `delta_t`, ...).  This module implements exactly the part of the
HDF5 File Format Specification (version 3) that libhdf5 emits for such a file with its default ("earliest") settings:

  superblock v0 -> root symbol-table entry -> group B-tree (v1, "TREE") + local heap ("HEAP") + symbol-table nodes ("SNOD")
  object headers v1 (+ continuation blocks): dataspace (v1 / v2), datatype (fixed, float, string, variable-length),
  data layout v3 (compact / contiguous; chunked without filters through the v1 chunk B-tree), attribute messages (v1 - v3)
  global heap collections ("GCOL") for the variable-length rows.

`File(path)[name]` gives a `Dataset` with `.shape`, `.dtype`, `len()`, integer / slice indexing returning numpy arrays (a
variable-length dataset yields one 1-D array per row), `File.attrs` a dict, `File.get(name, default)`, `name in File`.
`write_file` is the matching writer (same structures), used by the tests to make fixtures and by tools to export synthetic
data; this image has no h5py, so the reader is verified against this writer and the specification only -- with h5py present,
`dataloader.open_h5` uses h5py instead.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


# ======================================================================================================= reading
class _Type:
    """Decoded datatype message: numpy dtype of an element, or a variable-length sequence / string of a base type."""

    def __init__(self, kind: str, dtype: Optional[np.dtype], size: int, base: Optional["_Type"] = None):
        self.kind, self.dtype, self.size, self.base = kind, dtype, size, base      # kind: "num" | "str" | "vlen" | "vstr"


def _parse_type(b: bytes, o: int = 0) -> Tuple[_Type, int]:
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, o)
    cls, ver = cv & 15, cv >> 4
    o += 8
    if cls == 0:                                                   # fixed point
        order = ">" if (b0 & 1) else "<"
        signed = "i" if (b0 & 8) else "u"
        return _Type("num", np.dtype(f"{order}{signed}{size}"), size), o + 4
    if cls == 1:                                                   # floating point
        order = ">" if (b0 & 1) else "<"
        return _Type("num", np.dtype(f"{order}f{size}"), size), o + 12
    if cls == 3:                                                   # fixed-length string
        return _Type("str", np.dtype(f"S{size}"), size), o
    if cls == 9:                                                   # variable length: sequence (b0 & 15 == 0) or string (== 1)
        base, o2 = _parse_type(b, o)
        return _Type("vstr" if (b0 & 15) == 1 else "vlen", None, size, base), o2
    raise NotImplementedError(f"HDF5 datatype class {cls} (version {ver}) is outside the subset gen.py writes")


def _parse_space(b: bytes) -> Tuple[int, ...]:
    ver, rank, flags = struct.unpack_from("<BBB", b, 0)
    o = 8 if ver == 1 else 4
    return tuple(struct.unpack_from(f"<{rank}Q", b, o)) if rank else ()


class Dataset:
    def __init__(self, f: "File", name: str, msgs: List[Tuple[int, bytes]]):
        self._f, self.name = f, name
        self.attrs: Dict[str, object] = {}
        self._type: Optional[_Type] = None
        self.shape: Tuple[int, ...] = ()
        self._layout = None
        for t, d in msgs:
            if t == 0x1:
                self.shape = _parse_space(d)
            elif t == 0x3:
                self._type = _parse_type(d)[0]
            elif t == 0x8:
                self._layout = d
            elif t == 0xC:
                k, v = f._parse_attr(d)
                self.attrs[k] = v
        if self._type is None or self._layout is None:
            raise ValueError(f"HDF5 object {name!r} is not a dataset")
        self._raw: Optional[bytes] = None

    @property
    def dtype(self):
        t = self._type
        return t.dtype if t.kind in ("num", "str") else np.dtype("O")

    def __len__(self) -> int:
        return int(self.shape[0]) if self.shape else 1

    def _bytes(self) -> bytes:
        if self._raw is not None:
            return self._raw
        d, f = self._layout, self._f
        n = int(np.prod(self.shape)) if self.shape else 1
        total = n * self._type.size
        ver, cls = d[0], d[1]
        if ver != 3:
            raise NotImplementedError(f"HDF5 data layout message version {ver}")
        if cls == 0:                                               # compact
            sz = struct.unpack_from("<H", d, 2)[0]
            raw = d[4:4 + sz]
        elif cls == 1:                                             # contiguous
            addr, sz = struct.unpack_from("<QQ", d, 2)
            raw = b"\0" * total if addr == UNDEF else f._read(addr, sz)
        elif cls == 2:                                             # chunked, v1 B-tree, no filters
            rank = d[2]
            bt = struct.unpack_from("<Q", d, 3)[0]
            cdims = struct.unpack_from(f"<{rank}I", d, 11)
            raw = f._read_chunked(bt, self.shape, cdims[:-1], self._type.size)
        else:
            raise NotImplementedError(f"HDF5 layout class {cls}")
        self._raw = raw[:total]
        return self._raw

    def _row(self, i: int):
        t = self._type
        raw = self._bytes()
        inner = int(np.prod(self.shape[1:])) if len(self.shape) > 1 else 1
        if t.kind == "num" or t.kind == "str":
            a = np.frombuffer(raw, dtype=t.dtype, count=inner, offset=i * inner * t.size)
            a = a.reshape(self.shape[1:]) if len(self.shape) > 1 else a[0]
            return a
        out = []
        for j in range(inner):
            cnt, addr, idx = struct.unpack_from("<IQI", raw, (i * inner + j) * 16)
            obj = self._f._heap_object(addr, idx) if cnt else b""
            if t.kind == "vstr":
                out.append(obj[:cnt].decode("utf-8", "replace"))
            else:
                out.append(np.frombuffer(obj, dtype=t.base.dtype, count=cnt).copy())
        return out[0] if len(self.shape) <= 1 else out

    def __getitem__(self, key):
        if isinstance(key, tuple) and key == ():
            key = Ellipsis
        if key is Ellipsis:
            if not self.shape:
                return self._row(0)
            if self._type.kind == "num":
                return np.frombuffer(self._bytes(), dtype=self._type.dtype).reshape(self.shape).copy()
            key = slice(None)
        n = len(self)
        if isinstance(key, (int, np.integer)):
            i = int(key)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError(key)
            r = self._row(i)
            return r.copy() if isinstance(r, np.ndarray) else r
        if isinstance(key, slice):
            rows = [self._row(i) for i in range(*key.indices(n))]
            if self._type.kind == "num":
                return np.array(rows, dtype=self._type.dtype)
            arr = np.empty(len(rows), dtype=object)
            for i, r in enumerate(rows):
                arr[i] = r
            return arr
        raise TypeError(f"unsupported HDF5 index {key!r}")


class File:
    def __init__(self, path: str, mode: str = "r", **_ignored):
        if mode != "r":
            raise ValueError("gwb200 _hdf5.File is read-only (use write_file)")
        self.filename = path
        self._fh = open(path, "rb")
        head = self._read(0, 96)
        if head[:8] != SIG:
            raise OSError(f"{path}: not an HDF5 file")
        ver = head[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {ver}: written with a newer file-format setting than gen.py uses")
        if head[13] != 8 or head[14] != 8:
            raise NotImplementedError("HDF5 files with offsets / lengths other than 8 bytes")
        o = 24 + (4 if ver == 1 else 0)
        self._base = struct.unpack_from("<Q", head, o)[0]
        root = o + 32                                              # base, free-space, end-of-file, driver-info addresses
        _, hdr, cache = struct.unpack_from("<QQI", head, root)
        self._gcols: Dict[int, Dict[int, bytes]] = {}
        msgs = self._object_header(hdr)
        self.attrs: Dict[str, object] = {}
        self._links: Dict[str, int] = {}
        for t, d in msgs:
            if t == 0x11:
                bt, heap = struct.unpack_from("<QQ", d, 0)
                self._walk_group(bt, self._local_heap(heap))
            elif t == 0xC:
                try:
                    k, v = self._parse_attr(d)
                except (NotImplementedError, struct.error, ValueError, KeyError):
                    continue                                   # an attribute of a type outside the subset must not make the file unreadable
                self.attrs[k] = v
        self._cache: Dict[str, Dataset] = {}

    # ---- low level
    def _read(self, addr: int, n: int) -> bytes:
        self._fh.seek(addr + getattr(self, "_base", 0))
        return self._fh.read(n)

    def _object_header(self, addr: int) -> List[Tuple[int, bytes]]:
        ver, _, nmsg, _, size = struct.unpack("<BBHII", self._read(addr, 12))
        if ver != 1:
            raise NotImplementedError(f"HDF5 object header version {ver} (new-style): outside the subset gen.py's h5py defaults write")
        blocks = [(addr + 16, size)]
        msgs: List[Tuple[int, bytes]] = []
        while blocks and len(msgs) < nmsg:
            a, sz = blocks.pop(0)
            buf = self._read(a, sz)
            o = 0
            while o + 8 <= sz and len(msgs) < nmsg:
                t, ms, _fl = struct.unpack_from("<HHB", buf, o)
                d = buf[o + 8:o + 8 + ms]
                o += 8 + ms
                if t == 0x10:                                      # continuation
                    blocks.append(struct.unpack_from("<QQ", d, 0))
                msgs.append((t, d))
        return msgs

    def _local_heap(self, addr: int) -> bytes:
        h = self._read(addr, 32)
        if h[:4] != b"HEAP":
            raise OSError("HDF5: bad local heap signature")
        size, _free, data = struct.unpack_from("<QQQ", h, 8)
        return self._read(data, size)

    def _walk_group(self, addr: int, heap: bytes) -> None:
        h = self._read(addr, 24)
        if h[:4] == b"SNOD":
            n = struct.unpack_from("<H", h, 6)[0]
            ents = self._read(addr + 8, n * 40)
            for i in range(n):
                name_off, hdr = struct.unpack_from("<QQ", ents, i * 40)
                end = heap.index(b"\0", name_off)
                self._links[heap[name_off:end].decode()] = hdr
            return
        if h[:4] != b"TREE":
            raise OSError("HDF5: bad group B-tree signature")
        n = struct.unpack_from("<H", h, 6)[0]
        body = self._read(addr + 24, (2 * n + 1) * 8)
        for i in range(n):
            self._walk_group(struct.unpack_from("<Q", body, (2 * i + 1) * 8)[0], heap)

    def _read_chunked(self, bt: int, shape, cdims, esize: int) -> bytes:
        rank = len(shape)
        out = np.zeros(shape, dtype=np.uint8 if esize == 1 else np.dtype(f"V{esize}"))
        def walk(addr):
            h = self._read(addr, 24)
            if h[:4] != b"TREE":
                raise OSError("HDF5: bad chunk B-tree signature")
            level, n = h[5], struct.unpack_from("<H", h, 6)[0]
            ksz = 8 + 8 * (rank + 1)
            body = self._read(addr + 24, n * (ksz + 8) + ksz)
            for i in range(n):
                ko = i * (ksz + 8)
                csize, mask = struct.unpack_from("<II", body, ko)
                offs = struct.unpack_from(f"<{rank + 1}Q", body, ko + 8)[:rank]
                child = struct.unpack_from("<Q", body, ko + ksz)[0]
                if level > 0:
                    walk(child)
                    continue
                if mask:
                    raise NotImplementedError("HDF5 chunk filters (compression) are outside the subset gen.py writes")
                blk = np.frombuffer(self._read(child, csize), dtype=out.dtype).reshape(cdims)
                sl = tuple(slice(o_, min(o_ + c, s)) for o_, c, s in zip(offs, cdims, shape))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        if bt != UNDEF:
            walk(bt)
        return out.tobytes()

    def _heap_object(self, addr: int, idx: int) -> bytes:
        col = self._gcols.get(addr)
        if col is None:
            h = self._read(addr, 16)
            if h[:4] != b"GCOL":
                raise OSError("HDF5: bad global heap signature")
            size = struct.unpack_from("<Q", h, 8)[0]
            buf = self._read(addr, size)
            col, o = {}, 16
            while o + 16 <= size:
                i, _rc, _r, osz = struct.unpack_from("<HHIQ", buf, o)
                if i == 0:
                    break
                col[i] = buf[o + 16:o + 16 + osz]
                o += 16 + (osz + 7) // 8 * 8
            self._gcols[addr] = col
        return col[idx]

    def _parse_attr(self, d: bytes):
        ver = d[0]
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        o = 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = d[o:o + nsz].split(b"\0")[0].decode()
        o += pad(nsz)
        t, _ = _parse_type(d, o)
        o += pad(tsz)
        shape = _parse_space(d[o:o + ssz]) if ssz else ()
        o += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        if t.kind == "num":
            a = np.frombuffer(d, dtype=t.dtype, count=n, offset=o)
            return name, (a[0] if not shape else a.reshape(shape).copy())
        if t.kind == "str":
            s = d[o:o + t.size].split(b"\0")[0].decode("utf-8", "replace")
            return name, s
        if t.kind == "vstr":
            cnt, addr, idx = struct.unpack_from("<IQI", d, o)
            return name, (self._heap_object(addr, idx)[:cnt].decode("utf-8", "replace") if cnt else "")
        cnt, addr, idx = struct.unpack_from("<IQI", d, o)
        return name, np.frombuffer(self._heap_object(addr, idx), dtype=t.base.dtype, count=cnt).copy()

    # ---- h5py-like surface
    def keys(self):
        return self._links.keys()

    def __contains__(self, name: str) -> bool:
        return name in self._links

    def __getitem__(self, name: str) -> Dataset:
        ds = self._cache.get(name)
        if ds is None:
            if name not in self._links:
                raise KeyError(name)
            ds = Dataset(self, name, self._object_header(self._links[name]))
            self._cache[name] = ds
        return ds

    def get(self, name: str, default=None):
        return self[name] if name in self._links else default

    def close(self) -> None:
        try:
            self._fh.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ======================================================================================================= writing
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _type_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        exp, man, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        return struct.pack("<BBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize) + \
            struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man, exp, 0, man, bias)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)
    raise TypeError(dt)


def _vlen_type_msg(base: np.dtype) -> bytes:
    return struct.pack("<BBBBI", 0x19, 0, 0, 0, 16) + _type_msg(base)


def _space_msg(shape) -> bytes:
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _msg(t: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHBBBB", t, len(data), 0, 0, 0, 0) + data


def _attr_msg(name: str, value) -> bytes:
    nb = name.encode() + b"\0"
    if isinstance(value, str):
        vb = value.encode() + b"\0"
        tm, sm, data = _type_msg(np.dtype(f"S{len(vb)}")), _space_msg(()), vb
    else:
        a = np.asarray(value)
        if a.dtype.kind == "b":
            a = a.astype(np.int8)
        tm, sm, data = _type_msg(a.dtype), _space_msg(a.shape), a.tobytes()
    body = struct.pack("<BBHHH", 1, 0, len(nb), len(tm), len(sm)) + _pad8(nb) + _pad8(tm) + _pad8(sm) + data
    return _msg(0xC, body)


def _object_header(msgs: List[bytes]) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body


def write_file(path: str, datasets: Dict[str, object], attrs: Optional[Dict[str, object]] = None) -> None:
    """Write `datasets` (name -> numpy array, or a list of 1-D arrays = a variable-length dataset) and root `attrs` with the
    structures libhdf5 uses by default: superblock v0, symbol-table root group, v1 object headers, contiguous layout, a global
    heap collection per variable-length dataset."""
    attrs = attrs or {}
    names = sorted(datasets)                                       # the group B-tree keeps names in order
    blob = bytearray(b"\0" * 96)                                   # superblock placeholder

    def put(b: bytes) -> int:
        blob.extend(b"\0" * (-len(blob) % 8))
        a = len(blob)
        blob.extend(b)
        return a

    hdr_addr: Dict[str, int] = {}
    for name in names:
        v = datasets[name]
        if isinstance(v, (list, tuple)) or (isinstance(v, np.ndarray) and v.dtype == object):
            rows = [np.ascontiguousarray(r) for r in v]
            base = rows[0].dtype if rows else np.dtype("<f4")
            objs = b""
            refs = []
            for i, r in enumerate(rows):
                rb = r.astype(base).tobytes()
                objs += struct.pack("<HHIQ", i + 1, 1, 0, len(rb)) + _pad8(rb)
                refs.append((len(r), i + 1))
            size = 16 + len(objs) + 16
            size = max(4096, (size + 7) // 8 * 8)
            free = size - 16 - len(objs) - 16
            col = b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size) + objs + struct.pack("<HHIQ", 0, 0, 0, free)
            col = col + b"\0" * (size - len(col))
            gaddr = put(col)
            raw = b"".join(struct.pack("<IQI", n, gaddr, idx) for n, idx in refs)
            tm, shape = _vlen_type_msg(base), (len(rows),)
        else:
            a = np.ascontiguousarray(v)
            raw, tm, shape = a.tobytes(), _type_msg(a.dtype), a.shape
        daddr = put(raw) if raw else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, daddr, len(raw))
        hdr_addr[name] = put(_object_header([_msg(0x1, _space_msg(shape)), _msg(0x3, tm), _msg(0x8, layout)]))
    # local heap with the link names (offset 0 is the empty string), symbol-table nodes of <= 8 entries, one B-tree node
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for name in names:
        name_off[name] = len(heap)
        heap.extend(_pad8(name.encode() + b"\0"))
    heap_data = put(bytes(heap))
    heap_addr = put(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), UNDEF, heap_data))
    snods, keys = [], [0]
    for i in range(0, max(len(names), 1), 8):
        part = names[i:i + 8]
        ents = b"".join(struct.pack("<QQII16s", name_off[n], hdr_addr[n], 0, 0, b"") for n in part)
        ents += b"\0" * (40 * (8 - len(part)))
        snods.append(put(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)) + ents))
        keys.append(name_off[part[-1]] if part else 0)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
    for i, a in enumerate(snods):
        tree += struct.pack("<QQ", keys[i], a)
    tree += struct.pack("<Q", keys[-1])
    tree += b"\0" * ((2 * 32 + 1) * 8 - (len(tree) - 24))          # room for 2K = 32 children (internal node K = 16)
    bt_addr = put(tree)
    root_msgs = [_msg(0x11, struct.pack("<QQ", bt_addr, heap_addr))] + [_attr_msg(k, v) for k, v in attrs.items()]
    root_hdr = put(_object_header(root_msgs))
    eof = len(blob) + (-len(blob) % 8)
    blob.extend(b"\0" * (eof - len(blob)))
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", bt_addr, heap_addr)
    blob[:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(blob))
