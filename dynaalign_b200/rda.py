"""Reader for R ``.rda`` / ``.RData`` files (the reference ships its datasets as data/*.rda).

Standard-library only (gzip/bz2/lzma + struct).  Implements the subset of R's XDR serialization
format (versions 2 and 3: ``RDX2`` / ``RDX3`` headers) needed for data frames of character, integer,
double and logical columns, including attributes, references, and the ALTREP compact integer /
real sequences and deferred-string wrappers that R >= 3.5 uses for row names.

    objs = load_rda("data/evp_peparray.rda")          # {"evp_peparray": {"PROBE_SEQUENCE": [...], ...}}

Data frames come back as ``dict`` column -> list; factors as their integer codes with a
``levels`` entry under ``"<column>.levels"``.  This is the on-disk-format row of SURVEY.md section 8(f).
"""
import bz2
import gzip
import lzma
import struct

NILVALUE, REFSXP, ALTREP = 254, 255, 238
GLOBALENV, EMPTYENV, BASEENV, MISSINGARG, UNBOUND, BASENAMESPACE = 253, 242, 241, 251, 252, 247
SYMSXP, LISTSXP, CLOSXP, ENVSXP, LANGSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP, VECSXP, EXPRSXP, RAWSXP = \
    1, 2, 3, 4, 6, 9, 10, 13, 14, 15, 16, 19, 20, 24
ATTRLISTSXP, ATTRLANGSXP, NAMESPACESXP, PACKAGESXP, PERSISTSXP = 239, 240, 249, 250, 248

NA_INTEGER = -2147483648


class RObj:
    __slots__ = ("value", "attrs")

    def __init__(self, value, attrs=None):
        self.value = value
        self.attrs = attrs or {}


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def int(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self):
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def bytes(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE:
            return None
        if t in (GLOBALENV, EMPTYENV, BASEENV, MISSINGARG, UNBOUND, BASENAMESPACE):
            return RObj(("env", t))
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()
            o = RObj(("sym", name.value if isinstance(name, RObj) else name))
            self.refs.append(o)
            return o
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.int()
            n = self.int()
            o = RObj(("ns", [self.item() for _ in range(n)]))
            self.refs.append(o)
            return o
        if t == ENVSXP:
            o = RObj(("envir", None))
            self.refs.append(o)
            self.int()
            for _ in range(4):
                self.item()
            return o
        if t in (LISTSXP, LANGSXP, CLOSXP, ATTRLISTSXP, ATTRLANGSXP):
            # pairlist: iterate instead of recursing on the tail
            items = []
            attrs = None
            while True:
                if t in (ATTRLISTSXP, ATTRLANGSXP):
                    has_attr = True
                if has_attr:
                    attrs = self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                tagname = tag.value[1] if isinstance(tag, RObj) and isinstance(tag.value, tuple) else None
                items.append((tagname, car))
                flags = self.int()
                t = flags & 0xFF
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
                if t not in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                    self.p -= 4
                    tail = self.item()
                    if tail is not None:
                        items.append((None, tail))
                    break
            return RObj(("pairlist", items))
        if t == ALTREP:
            info = self.item()
            state = self.item()
            self.item()  # attributes
            cls = info.value[1][0][1].value[1] if isinstance(info, RObj) else ""
            return self._altrep(cls, state)
        if t == CHARSXP:
            n = self.int()
            if n == -1:
                return RObj(None)
            raw = self.bytes(n)
            enc = "latin-1" if flags & (1 << 14) else "utf-8"
            try:
                return RObj(raw.decode(enc))
            except UnicodeDecodeError:
                return RObj(raw.decode("latin-1"))
        if t == LGLSXP or t == INTSXP:
            n = self.length()
            v = list(struct.unpack_from(">%di" % n, self.b, self.p))
            self.p += 4 * n
            o = RObj(v)
        elif t == REALSXP:
            n = self.length()
            v = list(struct.unpack_from(">%dd" % n, self.b, self.p))
            self.p += 8 * n
            o = RObj(v)
        elif t == CPLXSXP:
            n = self.length()
            v = list(struct.unpack_from(">%dd" % (2 * n), self.b, self.p))
            self.p += 16 * n
            o = RObj(v)
        elif t == RAWSXP:
            n = self.length()
            o = RObj(self.bytes(n))
        elif t == STRSXP:
            n = self.length()
            o = RObj([self.item().value for _ in range(n)])
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            o = RObj([self.item() for _ in range(n)])
        else:
            raise ValueError("unsupported SEXP type %d at byte %d" % (t, self.p))
        if has_attr:
            a = self.item()
            if isinstance(a, RObj) and isinstance(a.value, tuple) and a.value[0] == "pairlist":
                o.attrs = {k: v for k, v in a.value[1]}
        return o

    def _altrep(self, cls, state):
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = state.value[:3]
            n = int(n)
            if cls == "compact_intseq":
                return RObj([int(start + i * step) for i in range(n)])
            return RObj([start + i * step for i in range(n)])
        if cls == "deferred_string":
            # state = pairlist(arg, scalar); arg is the numeric vector to be coerced
            arg = state.value[1][0][1]
            return RObj([str(int(x)) if float(x).is_integer() else repr(x) for x in arg.value])
        if cls.startswith("wrap_"):
            inner = state.value[0] if isinstance(state.value, list) else state.value[1][0][1]
            return inner
        raise ValueError("unsupported ALTREP class %r" % cls)


def _decompress(path):
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def _simplify(o):
    if o is None:
        return None
    if not isinstance(o, RObj):
        return o
    cls = o.attrs.get("class")
    names = o.attrs.get("names")
    if isinstance(o.value, list) and o.value and isinstance(o.value[0], RObj) or (names is not None and isinstance(o.value, list)
                                                                                  and all(isinstance(x, (RObj, type(None))) for x in o.value)):
        cols = [_simplify(x) for x in o.value]
        if names is not None:
            out = {}
            for nm, col, rawcol in zip(names.value, cols, o.value):
                out[nm] = col
                if isinstance(rawcol, RObj) and "levels" in rawcol.attrs:
                    out[nm + ".levels"] = rawcol.attrs["levels"].value
            return out
        return cols
    return o.value


def load_rda(path):
    """Parse an .rda file; returns {object name: value} (data frames as dict of columns)."""
    buf = _decompress(path)
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("not an RDX2/RDX3 file: %r" % buf[:5])
    r = _Reader(buf)
    r.p = 5
    if r.bytes(2) != b"X\n":
        raise ValueError("only XDR serialization is supported")
    version = r.int()
    r.int()
    r.int()
    if version == 3:
        n = r.int()
        r.bytes(n)
    top = r.item()
    out = {}
    for name, val in top.value[1]:
        out[name] = _simplify(val)
    return out


def load_sequences(path, column, obj=None):
    """Convenience: the character column ``column`` of the (first / named) data frame in ``path``."""
    objs = load_rda(path)
    df = objs[obj] if obj else next(iter(objs.values()))
    return list(df[column])
