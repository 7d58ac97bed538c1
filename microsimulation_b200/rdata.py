"""Reader for R's `.rda` / `.RData` files (gzip, bzip2 or xz around the XDR serialisation, "RDX2"/"RDX3"), in pure
Python: enough of the format for data sets made of lists, data frames, numeric / integer / logical / character vectors
and factors — what `data/fhcrcData.rda` of the reference holds (ten look-up tables of the FHCRC prostate model,
man/Data.Rd).  No R needed.

    from microsimulation_b200.rdata import read_rda
    fhcrc = read_rda("data/fhcrcData.rda")["fhcrcData"]      # dict of ten tables, each a dict of numpy columns

The serialisation (R internals, "Serialization formats"): every item starts with a 32-bit flags word — type in the low
byte, bits 8-10 "is object", "has attributes", "has tag" — followed by the payload; pairlists carry a symbol tag per
element, symbols and environments are back-referenced by index (REFSXP).
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE_SXP, REFSXP, NILSXP, SYMSXP, LISTSXP, CHARSXP = 254, 255, 0, 1, 2, 9
LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, LANGSXP = 10, 13, 14, 16, 19, 6
ALTREP_SXP, ATTRLISTSXP, ATTRLANGSXP = 238, 239, 240
GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP, BASENAMESPACE_SXP = 253, 242, 241, 251, 252, 247
NA_INT = -2147483648


class RVector(np.ndarray):
    """A numpy array that remembers its R attributes (names, levels, class, ...)."""

    def __new__(cls, arr, attrs=None):
        obj = np.asarray(arr).view(cls)
        obj.attrs = attrs or {}
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", {})


class RList(list):
    def __init__(self, items=(), attrs=None):
        super().__init__(items)
        self.attrs = attrs or {}


class _Reader:
    def __init__(self, data):
        self.b, self.p, self.refs = data, 0, []

    def take(self, n):
        out = self.b[self.p:self.p + n]
        if len(out) != n:
            raise ValueError("rda: unexpected end of data")
        self.p += n
        return out

    def i32(self):
        return struct.unpack(">i", self.take(4))[0]

    def length(self):
        n = self.i32()
        if n == -1:  # long vector: two more words
            hi, lo = struct.unpack(">II", self.take(8))
            n = (hi << 32) | lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xff
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t in (NILVALUE_SXP, NILSXP):
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP, BASENAMESPACE_SXP):
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()  # a CHARSXP
            self.refs.append(name)
            return name
        if t == CHARSXP:
            n = self.i32()
            if n == -1:
                return None  # NA_character_
            raw = self.take(n)
            return raw.decode("latin-1" if flags & (1 << 14) else "utf-8", errors="replace")
        if t in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):  # pairlist: (attributes) (tag) car cdr, cdr iterated
            out = []
            while True:
                if t in (ATTRLISTSXP, ATTRLANGSXP) or has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                out.append((tag, self.item()))
                flags = self.i32()
                t = flags & 0xff
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
                if t in (NILVALUE_SXP, NILSXP):
                    return out
                if t not in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                    raise ValueError("rda: malformed pairlist (type %d)" % t)
        if t in (LGLSXP, INTSXP):
            n = self.length()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
            return self._finish(v, has_attr, t)
        if t == REALSXP:
            n = self.length()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
            return self._finish(v, has_attr, t)
        if t == STRSXP:
            n = self.length()
            v = np.array([self.item() for _ in range(n)], dtype=object)
            return self._finish(v, has_attr, t)
        if t == VECSXP:
            n = self.length()
            items = [self.item() for _ in range(n)]
            attrs = dict(self.item() or []) if has_attr else {}
            return RList(items, attrs)
        if t == ALTREP_SXP:  # info pairlist, state, attributes: compact sequences and wrapped vectors
            info, state, attr = self.item(), self.item(), self.item()
            cls = info[0][1] if info else None
            if cls == "compact_intseq" or cls == "compact_realseq":
                n, start, step = [float(x) for x in state[:3]]
                return RVector(start + step * np.arange(int(n)), dict(attr or []))
            if cls in ("wrap_real", "wrap_integer", "wrap_logical", "wrap_string"):
                inner = state[0] if isinstance(state, RList) else state
                return inner
            raise ValueError("rda: unsupported ALTREP class %r" % (cls,))
        raise ValueError("rda: unsupported SEXP type %d" % t)

    def _finish(self, v, has_attr, t):
        attrs = dict(self.item() or []) if has_attr else {}
        return RVector(v, attrs)


def _simplify(x):
    """R objects -> plain Python: data frames and named lists become dicts, factors become their labels."""
    if isinstance(x, RList):
        names = x.attrs.get("names")
        vals = [_simplify(v) for v in x]
        if names is not None and len(names) == len(vals):
            return {str(k): v for k, v in zip(names, vals)}
        return vals
    if isinstance(x, RVector):
        levels = x.attrs.get("levels")
        if levels is not None and x.dtype == np.int32:  # factor: 1-based codes into levels
            lv = np.asarray(levels, dtype=object)
            return np.array([None if c == NA_INT else lv[c - 1] for c in np.asarray(x)], dtype=object)
        a = np.asarray(x)
        if a.dtype == np.int32 and (a == NA_INT).any():
            a = np.where(a == NA_INT, np.nan, a.astype(np.float64))
        return a
    return x


def read_rda(path, simplify=True):
    """All objects saved in an .rda file, by name."""
    raw = open(path, "rb").read()
    for opener in (gzip.decompress, bz2.decompress, lzma.decompress, lambda b: b):
        try:
            data = opener(raw)
            if data[:3] == b"RDX":
                break
        except Exception:
            continue
    else:
        raise ValueError("%s: not an R data file" % path)
    if data[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("%s: unsupported R data format %r" % (path, data[:5]))
    r = _Reader(data)
    r.p = 5
    if r.take(2) != b"X\n":
        raise ValueError("%s: only the XDR serialisation is supported" % path)
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:  # native encoding name
        r.take(r.i32())
    top = r.item()  # a pairlist of (name, value)
    out = {}
    for tag, val in top or []:
        out[str(tag)] = _simplify(val) if simplify else val
    return out
