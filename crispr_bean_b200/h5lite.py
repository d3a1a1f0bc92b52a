"""Minimal pure-Python reader for the HDF5 subset `.h5ad` files written by anndata / h5py use.

Neither h5py nor anndata is available where this package is built and tested, and `bean run` takes its input as an
`.h5ad` ReporterScreen (bean/framework/ReporterScreen.py:1009 `read_h5ad`).  The files h5py
writes by default use the original HDF5 structures only, which is what this module understands:

  superblock version 0, version-1 object headers (with continuation blocks), symbol-table groups (v1 B-tree "TREE" +
  "SNOD" nodes + local "HEAP"), contiguous / compact / chunked (v1 B-tree) dataset layouts with the deflate and shuffle
  filters, fixed-point, floating-point, fixed-length string, enum (h5py bool) and variable-length string datatypes
  (global heap "GCOL"), attributes (message versions 1-3).

Written from the HDF5 File Format Specification (version 1.1 structures); nothing of libhdf5 / h5py is used.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Optional

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Datatype:
    """Parsed datatype message: numpy dtype for plain classes, flags for strings / vlen strings / enums."""

    def __init__(self, buf: bytes, off: int = 0):
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
        self.cls, self.version, self.size = cv & 0x0F, cv >> 4, size
        self.vlen_str = False
        self.np: Optional[np.dtype] = None
        self.enum_names: Optional[list] = None
        props = off + 8
        if self.cls == 0:  # fixed point
            order = ">" if b0 & 1 else "<"
            signed = bool(b0 & 0x08)
            self.np = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
            self.end = props + 4
        elif self.cls == 1:  # floating point
            order = ">" if b0 & 1 else "<"
            self.np = np.dtype(f"{order}f{size}")
            self.end = props + 12
        elif self.cls == 3:  # fixed-length string
            self.np = np.dtype(f"S{size}")
            self.end = props
        elif self.cls == 9:  # variable length: string when the type bits say so, else a sequence (unsupported)
            self.vlen_str = (b0 & 0x0F) == 1
            base = _Datatype(buf, props)
            self.base, self.end = base, base.end
            if not self.vlen_str:
                raise H5Error("variable-length sequences are not supported")
        elif self.cls == 8:  # enumeration (h5py stores bool as an enum of int8)
            n = b0 | (b1 << 8)
            base = _Datatype(buf, props)
            p = base.end
            names = []
            for _ in range(n):
                e = buf.index(b"\x00", p)
                names.append(buf[p:e].decode())
                p += _pad8(e - p + 1) if self.version < 3 else e - p + 1
            p += n * base.size
            self.np, self.enum_names, self.end = base.np, names, p
        elif self.cls == 6:  # compound (old anndata data frames): parsed member by member
            n = b0 | (b1 << 8)
            p = props
            names, formats, offsets = [], [], []
            for _ in range(n):
                e = buf.index(b"\x00", p)
                names.append(buf[p:e].decode())
                p += _pad8(e - p + 1) if self.version < 3 else e - p + 1
                if self.version == 1:
                    (moff,) = struct.unpack_from("<I", buf, p)
                    p += 4 + 1 + 3 + 4 + 4 + 16
                elif self.version == 2:
                    (moff,) = struct.unpack_from("<I", buf, p)
                    p += 4
                else:
                    nb = max(1, (size.bit_length() + 7) // 8)
                    moff = int.from_bytes(buf[p:p + nb], "little")
                    p += nb
                m = _Datatype(buf, p)
                p = m.end
                names_ok = m.np is not None
                if not names_ok:
                    raise H5Error("compound member of unsupported type")
                formats.append(m.np)
                offsets.append(moff)
            self.np = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size})
            self.end = p
        else:
            raise H5Error(f"datatype class {self.cls} is not supported")


class _Object:
    """One object header: its messages decoded into what a reader needs."""

    def __init__(self, f: "File", addr: int):
        self.f, self.addr = f, addr
        self.dtype: Optional[_Datatype] = None
        self.shape: Optional[tuple] = None
        self.layout = None
        self.filters: list = []
        self.attrs: Dict[str, object] = {}
        self.symtab = None  # (btree, heap) for groups
        self._parse()

    def _parse(self):
        b = self.f.buf
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, self.addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} is not supported (file written with libver='latest'?)")
        blocks = [(self.addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                data = p + 8
                self._message(mtype, data, msize, blocks)
                seen += 1
                p = data + msize

    def _message(self, mtype, p, size, blocks):
        b, O = self.f.buf, self.f.O
        if mtype == 0x0001:  # dataspace
            ver, rank, flags = struct.unpack_from("<BBB", b, p)
            q = p + (8 if ver == 1 else 4)
            self.shape = tuple(struct.unpack_from(f"<{rank}Q", b, q)) if rank else ()
            if ver == 2 and b[p + 3] == 2:
                self.shape = None  # null dataspace
        elif mtype == 0x0003:
            self.dtype = _Datatype(b, p)
        elif mtype == 0x0008:  # layout
            ver, cls = struct.unpack_from("<BB", b, p)
            if ver != 3:
                raise H5Error(f"data layout version {ver} is not supported")
            if cls == 0:
                (n,) = struct.unpack_from("<H", b, p + 2)
                self.layout = ("compact", p + 4, n)
            elif cls == 1:
                addr, n = struct.unpack_from("<QQ", b, p + 2)
                self.layout = ("contiguous", addr, n)
            else:
                nd = b[p + 2]
                (bt,) = struct.unpack_from("<Q", b, p + 3)
                dims = struct.unpack_from(f"<{nd}I", b, p + 3 + O)
                self.layout = ("chunked", bt, dims)
        elif mtype == 0x000B:  # filter pipeline
            ver, nf = struct.unpack_from("<BB", b, p)
            q = p + (8 if ver == 1 else 2)
            for _ in range(nf):
                fid, = struct.unpack_from("<H", b, q)
                if ver == 1 or fid >= 256:
                    nlen, _fl, ncd = struct.unpack_from("<HHH", b, q + 2)
                    q += 8 + (_pad8(nlen) if ver == 1 else nlen)
                else:
                    _fl, ncd = struct.unpack_from("<HH", b, q + 2)
                    q += 6
                cd = struct.unpack_from(f"<{ncd}I", b, q)
                q += 4 * ncd + (4 if (ver == 1 and ncd % 2) else 0)
                self.filters.append((fid, cd))
        elif mtype == 0x000C:  # attribute
            ver = b[p]
            if ver == 1:
                nsz, dsz, ssz = struct.unpack_from("<HHH", b, p + 2)
                q = p + 8
                name = b[q:q + nsz].split(b"\x00")[0].decode()
                q += _pad8(nsz)
                dt = _Datatype(b, q)
                q += _pad8(dsz)
                shape, null = self._space(q)
                q += _pad8(ssz)
            else:
                nsz, dsz, ssz = struct.unpack_from("<HHH", b, p + 2)
                q = p + 8 + (1 if ver == 3 else 0)
                name = b[q:q + nsz].split(b"\x00")[0].decode()
                q += nsz
                dt = _Datatype(b, q)
                q += dsz
                shape, null = self._space(q)
                q += ssz
            self.attrs[name] = None if null else self.f._decode(dt, shape, b, q)
        elif mtype == 0x0010:  # continuation
            off, ln = struct.unpack_from("<QQ", b, p)
            blocks.append((off, ln))
        elif mtype == 0x0011:  # symbol table
            self.symtab = struct.unpack_from("<QQ", b, p)

    def _space(self, q):
        b = self.f.buf
        ver, rank = b[q], b[q + 1]
        if ver == 2 and b[q + 3] == 2:
            return (), True
        off = q + (8 if ver == 1 else 4)
        return (tuple(struct.unpack_from(f"<{rank}Q", b, off)) if rank else ()), False

    # -- groups ---------------------------------------------------------------------------------------
    def children(self) -> Dict[str, int]:
        if self.symtab is None:
            return {}
        bt, heap = self.symtab
        b = self.f.buf
        if b[heap:heap + 4] != b"HEAP":
            raise H5Error("bad local heap")
        (data,) = struct.unpack_from("<Q", b, heap + 8 + 16)
        out: Dict[str, int] = {}

        def walk(addr):
            if b[addr:addr + 4] == b"TREE":
                ntype, level, n = struct.unpack_from("<BBH", b, addr + 4)
                p = addr + 8 + 16
                for i in range(n):
                    (child,) = struct.unpack_from("<Q", b, p + 8)
                    walk(child)
                    p += 16
            elif b[addr:addr + 4] == b"SNOD":
                (n,) = struct.unpack_from("<H", b, addr + 6)
                p = addr + 8
                for _ in range(n):
                    noff, oaddr = struct.unpack_from("<QQ", b, p)
                    e = b.index(b"\x00", data + noff)
                    out[b[data + noff:e].decode()] = oaddr
                    p += 40
            else:
                raise H5Error("bad group node")

        walk(bt)
        return out

    # -- datasets -------------------------------------------------------------------------------------
    def read(self):
        if self.dtype is None or self.shape is None or self.layout is None:
            raise H5Error("not a dataset")
        dt, b = self.dtype, self.f.buf
        esize = dt.size
        n = int(np.prod(self.shape)) if self.shape else 1
        kind = self.layout[0]
        if kind == "compact":
            raw = b[self.layout[1]:self.layout[1] + self.layout[2]]
        elif kind == "contiguous":
            addr = self.layout[1]
            raw = b"" if addr == UNDEF else b[addr:addr + n * esize]
            if addr == UNDEF:
                raw = bytes(n * esize)
        else:
            raw = self._read_chunked(esize)
        return self.f._decode(dt, self.shape, raw, 0)

    def _read_chunked(self, esize):
        b = self.f.buf
        _, bt, cdims = self.layout
        chunk = cdims[:-1]
        rank = len(chunk)
        shape = self.shape
        out = np.zeros(shape, dtype=np.dtype(f"V{esize}"))
        if bt == UNDEF:
            return out.tobytes()

        def walk(addr):
            if b[addr:addr + 4] != b"TREE":
                raise H5Error("bad chunk B-tree")
            ntype, level, n = struct.unpack_from("<BBH", b, addr + 4)
            p = addr + 8 + 16
            ksz = 8 + 8 * (rank + 1)
            for _ in range(n):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from(f"<{rank}Q", b, p + 8)
                (child,) = struct.unpack_from("<Q", b, p + ksz)
                if level > 0:
                    walk(child)
                else:
                    data = b[child:child + csize]
                    for i, (fid, cd) in reversed(list(enumerate(self.filters))):
                        if fmask & (1 << i):
                            continue
                        if fid == 1:
                            data = zlib.decompress(data)
                        elif fid == 2:  # shuffle
                            k = cd[0] if cd else esize
                            a = np.frombuffer(data, dtype=np.uint8)
                            data = a.reshape(k, -1).T.tobytes()
                        elif fid == 3:  # fletcher32: checksum trailer
                            data = data[:-4]
                        else:
                            raise H5Error(f"filter {fid} is not supported")
                    arr = np.frombuffer(data, dtype=np.dtype(f"V{esize}"), count=int(np.prod(chunk))).reshape(chunk)
                    sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                    sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                    out[sl_out] = arr[sl_in]
                p += ksz + 8

        walk(bt)
        return out.tobytes()


class File:
    """`File(path)[name]` -> Group / numpy array, `.attrs`, `.keys()`; the read-only subset of h5py's interface."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        if b[8] not in (0, 1):
            raise H5Error(f"superblock version {b[8]} is not supported (file written with libver='latest'?)")
        self.O, self.L = b[13], b[14]
        if (self.O, self.L) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = 24 + (4 if b[8] == 1 else 0)
        p += 4 * 8  # base address, free-space info, end of file, driver info
        _, root_addr = struct.unpack_from("<QQ", b, p)
        self._gcol: Dict[int, Dict[int, bytes]] = {}
        self.root = Group(self, _Object(self, root_addr), "/")

    def _heap_object(self, addr: int, idx: int) -> bytes:
        col = self._gcol.get(addr)
        if col is None:
            b = self.buf
            if b[addr:addr + 4] != b"GCOL":
                raise H5Error("bad global heap collection")
            (size,) = struct.unpack_from("<Q", b, addr + 8)
            col, p, end = {}, addr + 16, addr + size
            while p + 16 <= end:
                i, _rc, _, osz = struct.unpack_from("<HHIQ", b, p)
                if i == 0:
                    break
                col[i] = b[p + 16:p + 16 + osz]
                p += 16 + _pad8(osz)
            self._gcol[addr] = col
        return col.get(idx, b"")

    def _decode(self, dt: _Datatype, shape, raw, off):
        n = int(np.prod(shape)) if shape else 1
        if dt.vlen_str:
            out = np.empty(n, dtype=object)
            for i in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, off + 16 * i)
                out[i] = self._heap_object(addr, idx)[:ln].decode("utf-8", "replace") if addr not in (0, UNDEF) else ""
            return out.reshape(shape) if shape else out[0]
        arr = np.frombuffer(raw, dtype=dt.np, count=n, offset=off)
        if dt.enum_names is not None and sorted(dt.enum_names) == ["FALSE", "TRUE"]:
            arr = arr.astype(bool)
        if dt.np.kind == "S":
            arr = np.array([x.split(b"\x00")[0].decode("utf-8", "replace") for x in arr], dtype=object)
        arr = arr.reshape(shape) if shape else arr[0]
        return arr.copy() if isinstance(arr, np.ndarray) else arr

    def __getitem__(self, name):
        return self.root[name]

    def keys(self):
        return self.root.keys()

    @property
    def attrs(self):
        return self.root.attrs


class Group:
    def __init__(self, f: File, obj: _Object, name: str):
        self.f, self.obj, self.name = f, obj, name
        self._children = obj.children()

    @property
    def attrs(self):
        return self.obj.attrs

    def keys(self):
        return list(self._children)

    def __contains__(self, k):
        return k in self._children

    def __getitem__(self, name: str):
        node = self
        for part in [p for p in name.split("/") if p]:
            if not isinstance(node, Group) or part not in node._children:
                raise KeyError(name)
            obj = _Object(node.f, node._children[part])
            node = Group(node.f, obj, f"{node.name.rstrip('/')}/{part}") if obj.symtab is not None else Dataset(obj, part)
        return node


class Dataset:
    def __init__(self, obj: _Object, name: str):
        self.obj, self.name = obj, name
        self.shape, self.attrs = obj.shape, obj.attrs

    def __getitem__(self, key):
        a = self.obj.read()
        return a if key == () or key is Ellipsis else a[key]

    def read(self):
        return self.obj.read()


def tree(node, indent=0, out=None):
    """Text dump of a file's structure (debugging aid)."""
    out = [] if out is None else out
    if isinstance(node, File):
        node = node.root
    for k in node.keys():
        child = node[k]
        if isinstance(child, Group):
            out.append("  " * indent + f"{k}/  {dict(child.attrs)}")
            tree(child, indent + 1, out)
        else:
            dt = child.obj.dtype
            kind = "vlen-str" if dt.vlen_str else str(dt.np)
            out.append("  " * indent + f"{k}  shape={child.shape} {kind} layout={child.obj.layout[0]} attrs={dict(child.attrs)}")
    return out
