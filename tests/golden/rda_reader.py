"""Minimal reader for R's `.rda` (RDX2/RDX3, XDR) files — enough for the four packaged data files of
the reference (`/root/reference/data/*.rda`): numeric vectors/matrices, lists, raw vectors, character
vectors, attributes (dim, dimnames, names, class).  Used only by `make_golden.py` to turn the
reference's packaged data into the small `.npz` fixtures committed next to this file (there is no R
in the build image and `/root/reference` does not exist on the GPU box).
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE_SXP, SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, RAWSXP = (
    254, 1, 2, 9, 10, 13, 14, 16, 19, 24)
REFSXP, NAMESPACESXP, GLOBALENV_SXP, ALTREP_SXP, ATTRLISTSXP, LANGSXP = 255, 249, 253, 238, 239, 6


class RObj:
    def __init__(self, value, attrs=None):
        self.value = value
        self.attrs = attrs or {}

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attrs={list(self.attrs)})"


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs = []

    def int(self):
        (v,) = struct.unpack_from(">i", self.b, self.p)
        self.p += 4
        return v

    def bytes(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.int()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if typ == NILVALUE_SXP:
            return None
        if typ == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if typ == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if typ in (LISTSXP, LANGSXP, ATTRLISTSXP):
            out = []
            while True:
                attrs = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.int()
                typ2 = flags & 0xFF
                if typ2 == NILVALUE_SXP:
                    break
                if typ2 not in (LISTSXP, LANGSXP, ATTRLISTSXP):
                    raise ValueError(f"unexpected cdr type {typ2}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if typ == CHARSXP:
            n = self.int()
            return None if n == -1 else self.bytes(n).decode("latin1")
        if typ in (LGLSXP, INTSXP):
            n = self.int()
            v = np.frombuffer(self.bytes(4 * n), dtype=">i4").astype(np.int32)
        elif typ == REALSXP:
            n = self.int()
            v = np.frombuffer(self.bytes(8 * n), dtype=">f8").astype(np.float64)
        elif typ == RAWSXP:
            n = self.int()
            v = np.frombuffer(self.bytes(n), dtype=np.uint8).copy()
        elif typ == STRSXP:
            n = self.int()
            v = [self.item() for _ in range(n)]
        elif typ == VECSXP:
            n = self.int()
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {typ} at offset {self.p}")
        attrs = {}
        if has_attr:
            for tag, val in self.item():
                attrs[tag] = val
        return RObj(v, attrs)


def read_rda(path: str) -> dict:
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    if raw[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("not an RDX2/RDX3 file")
    version3 = raw[:5] == b"RDX3\n"
    if raw[5:7] != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    r = _Reader(raw)
    r.p = 7
    r.int(); r.int(); r.int()  # serialisation version, writer R version, minimal R version
    if version3:
        n = r.int()
        r.bytes(n)  # native encoding
    top = r.item()
    return {tag: val for tag, val in top}


def as_array(o: RObj) -> np.ndarray:
    """numeric/raw vector with a `dim` attribute -> numpy array in R's (column-major) index order."""
    v = np.asarray(o.value)
    dim = o.attrs.get("dim")
    if dim is not None:
        v = v.reshape(tuple(int(d) for d in dim.value), order="F")
    return v
