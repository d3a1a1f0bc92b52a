"""Pure-Python bigWig reader: per-base values of a genomic interval, which is all `bean run --acc-bw-path` asks of pyBigWig
(`track.values(chrom, start, end)`, bean/preprocessing/utils.py:92-104).  pyBigWig is not installable here.

Format (UCSC bigWig, Kent et al. 2010): 64-byte header, zoom headers, total summary, a B+ tree of chromosome names -> (id,
size), an R-tree over (chromosome id, base) rectangles pointing at data blocks, each block (zlib-compressed when
`uncompressBufSize` > 0) a 24-byte section header plus bedGraph / variableStep / fixedStep items.  Zoom levels are not
used: `values()` is exact by definition.
"""
from __future__ import annotations

import builtins
import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

BIGWIG_MAGIC, CHROM_TREE_MAGIC, RTREE_MAGIC = 0x888FFC26, 0x78CA8C91, 0x2468ACE0


class BigWigError(Exception):
    pass


class BigWig:
    def __init__(self, path: str):
        with builtins.open(path, "rb") as f:
            self._buf = f.read()
        b = self._buf
        if len(b) < 64:
            raise BigWigError("not a bigWig file")
        if struct.unpack_from("<I", b, 0)[0] == BIGWIG_MAGIC:
            self._e = "<"
        elif struct.unpack_from(">I", b, 0)[0] == BIGWIG_MAGIC:
            self._e = ">"
        else:
            raise BigWigError("not a bigWig file")
        (_, self.version, self.n_zoom, chrom_off, self._data_off, self._index_off, _, _, _, summary_off,
         self._uncompress, _) = self._u("IHHQQQHHQQIQ", 0)
        self.summary = None
        if summary_off:
            covered, lo, hi, total, sq = self._u("Qdddd", summary_off)
            self.summary = {"nBasesCovered": covered, "minVal": lo, "maxVal": hi, "sumData": total, "sumSquared": sq}
        self._chroms = self._read_chrom_tree(chrom_off)
        self._ids = {name: cid for name, (cid, _) in self._chroms.items()}

    def _u(self, fmt, off):
        return struct.unpack_from(self._e + fmt, self._buf, off)

    # ---- chromosome B+ tree ----------------------------------------------------------------------------------
    def _read_chrom_tree(self, off) -> Dict[str, Tuple[int, int]]:
        magic, _, key_size, val_size, _, _ = self._u("IIIIQQ", off)
        if magic != CHROM_TREE_MAGIC or val_size != 8:
            raise BigWigError("bad chromosome tree")
        out = {}

        def walk(p):
            is_leaf, _, count = self._u("BBH", p)
            p += 4
            for _ in range(count):
                key = self._buf[p:p + key_size].split(b"\0", 1)[0].decode()
                if is_leaf:
                    cid, size = self._u("II", p + key_size)
                    out[key] = (cid, size)
                else:
                    walk(self._u("Q", p + key_size)[0])
                p += key_size + 8

        walk(off + 32)
        return out

    def chroms(self) -> Dict[str, int]:
        return {k: size for k, (_, size) in self._chroms.items()}

    # ---- R-tree ---------------------------------------------------------------------------------------------
    def _blocks(self, cid: int, start: int, end: int) -> List[Tuple[int, int]]:
        if self._u("I", self._index_off)[0] != RTREE_MAGIC:
            raise BigWigError("bad R-tree index")
        found = []
        lo, hi = (cid, start), (cid, end)

        def walk(p):
            is_leaf, _, count = self._u("BBH", p)
            p += 4
            for _ in range(count):
                c0, b0, c1, b1 = self._u("IIII", p)
                overlaps = (c0, b0) < hi and (c1, b1) > lo
                if is_leaf:
                    if overlaps:
                        found.append(self._u("QQ", p + 16))
                    p += 32
                else:
                    if overlaps:
                        walk(self._u("Q", p + 16)[0])
                    p += 24

        walk(self._index_off + 48)
        return found

    # ---- data -----------------------------------------------------------------------------------------------
    def _intervals(self, off: int, size: int):
        """(chrom id, starts, ends, values) of one data block."""
        raw = self._buf[off:off + size]
        if self._uncompress:
            raw = zlib.decompress(raw)
        cid, c_start, _, step, span, kind, _, n = struct.unpack_from(self._e + "IIIIIBBH", raw, 0)
        body = raw[24:]
        e = self._e
        if kind == 1:  # bedGraph: start, end, value
            rec = np.frombuffer(body, dtype=np.dtype([("s", e + "u4"), ("e", e + "u4"), ("v", e + "f4")]), count=n)
            return cid, rec["s"].astype(np.int64), rec["e"].astype(np.int64), rec["v"]
        if kind == 2:  # variableStep: start, value (span from the header)
            rec = np.frombuffer(body, dtype=np.dtype([("s", e + "u4"), ("v", e + "f4")]), count=n)
            s = rec["s"].astype(np.int64)
            return cid, s, s + span, rec["v"]
        if kind == 3:  # fixedStep: values at chromStart + i * step
            v = np.frombuffer(body, dtype=e + "f4", count=n)
            s = c_start + step * np.arange(n, dtype=np.int64)
            return cid, s, s + span, v
        raise BigWigError(f"unknown section type {kind}")

    def values(self, chrom: str, start: int, end: int) -> np.ndarray:
        """Per-base signal of [start, end) as float64 holding the file's float32 values; NaN where the track has no data
        (pyBigWig.bigWigFile.values)."""
        if chrom not in self._ids:
            raise BigWigError(f"Invalid interval bounds! chromosome {chrom} is not in the file")
        cid, size = self._chroms[chrom]
        start, end = int(start), int(end)
        if not (0 <= start < end <= size):
            raise BigWigError("Invalid interval bounds!")
        out = np.full(end - start, np.nan, dtype=np.float32)
        for off, nbytes in self._blocks(cid, start, end):
            bcid, s, e, v = self._intervals(off, nbytes)
            if bcid != cid:
                continue
            keep = (e > start) & (s < end)
            for si, ei, vi in zip(np.maximum(s[keep], start) - start, np.minimum(e[keep], end) - start, v[keep]):
                out[si:ei] = vi
        return out.astype(np.float64)

    def coverage_summary(self) -> dict:
        """Covered bases, min, max and sum of the full-resolution data, recomputed from every block (the header's total
        summary states the same four numbers: a known answer the file carries about itself)."""
        covered, total, lo, hi = 0, 0.0, np.inf, -np.inf
        for cid_name, (cid, size) in self._chroms.items():
            for off, nbytes in self._blocks(cid, 0, size):
                bcid, s, e, v = self._intervals(off, nbytes)
                if bcid != cid or len(v) == 0:
                    continue
                w = (e - s).astype(np.float64)
                covered += int(w.sum())
                total += float((w * v.astype(np.float64)).sum())
                lo, hi = min(lo, float(v.min())), max(hi, float(v.max()))
        return {"nBasesCovered": covered, "minVal": lo, "maxVal": hi, "sumData": total}


def open(path: str) -> BigWig:  # noqa: A001  (pyBigWig.open)
    return BigWig(path)
