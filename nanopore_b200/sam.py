"""Pure-Python SAM reader/writer exposing the slice of the (old, 0.7-era) pysam API the realignment path uses.

pysam is absent here; the path touches only text SAM (`pysam.Samfile(path, "r")`, `pysam.Samfile(path, "wh",
template=sam)`, reference nanopore/analyses/utils.py:444,455,561,594-596) and these `AlignedRead` members:
qname, rname (reference id), pos, aend, qstart, qend, query, seq, cigar, cigarstring, is_reverse,
aligned_pairs, rnext (utils.py:168-180,295-386,487-506,557-609).  Coordinates are 0-based like pysam's.
"""
import re

import numpy as np

BAM_CMATCH, BAM_CINS, BAM_CDEL, BAM_CREF_SKIP, BAM_CSOFT_CLIP, BAM_CHARD_CLIP, BAM_CPAD, BAM_CEQUAL, BAM_CDIFF = range(9)
_OPS = "MIDNSHP=X"
_OP_CODE = {c: i for i, c in enumerate(_OPS)}
_CIGAR_TOKEN = re.compile(r"([0-9]+)([MIDNSHP=X])")
_REF_CONSUMING = (BAM_CMATCH, BAM_CDEL, BAM_CREF_SKIP, BAM_CEQUAL, BAM_CDIFF)
_QUERY_CONSUMING = (BAM_CMATCH, BAM_CINS, BAM_CEQUAL, BAM_CDIFF)


def parse_cigar(s):
    if s == "*" or s == "":
        return None
    toks = _CIGAR_TOKEN.findall(s)
    if sum(len(n) for n, _ in toks) + len(toks) != len(s):          # every character belongs to a <number><op> token
        raise ValueError("malformed CIGAR %r" % s)
    code = _OP_CODE
    return tuple([(code[op], int(n)) for n, op in toks])


def format_cigar(cigar):
    if not cigar:
        return "*"
    return "".join("%d%s" % (ln, _OPS[op]) for op, ln in cigar)


_OP_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(_OPS.encode()):
    _OP_LUT[_c] = _i
_OP_CHARS = np.array(list(_OPS))


def parse_cigar_arrays(s):
    """CIGAR text -> (op codes uint8, lengths int64) without building Python tuples (digits are combined with
    positional weights in numpy).  Same validation as parse_cigar."""
    if s == "*" or s == "":
        return np.zeros(0, np.uint8), np.zeros(0, np.int64)
    raw = np.frombuffer(s.encode("ascii"), dtype=np.uint8)
    code_at = _OP_LUT[raw]
    isop = code_at != 255
    isdig = (raw >= 48) & (raw <= 57)
    idx = np.flatnonzero(isop)
    if len(idx) == 0 or not (isop | isdig).all() or not isop[-1]:
        raise ValueError("malformed CIGAR %r" % s)
    starts = np.concatenate(([0], idx[:-1] + 1))
    if (idx == starts).any() or (idx - starts > 18).any():                       # an op without a number / absurd length
        raise ValueError("malformed CIGAR %r" % s)
    grp = np.cumsum(isop) - isop
    expo = idx[grp] - 1 - np.arange(len(raw))
    expo[isop] = 0
    val = np.where(isop, 0, (raw.astype(np.int64) - 48) * 10 ** expo)
    return code_at[idx], np.add.reduceat(val, starts)


def format_cigar_arrays(codes, lens):
    """(op codes, lengths) -> CIGAR text."""
    if len(codes) == 0:
        return "*"
    parts = np.char.add(np.asarray(lens, dtype=np.int64).astype("U20"), _OP_CHARS[np.asarray(codes, dtype=np.int64)])
    return "".join(parts.tolist())


class AlignedRead:
    def __init__(self):
        self.qname = ""
        self.flag = 0
        self.rname = -1
        self.pos = -1
        self.mapq = 0
        self._cigar = None         # tuple of (op, length), parsed on first use
        self._cigar_s = None       # CIGAR text as read from the file (until .cigar is assigned)
        self._cigar_a = None       # (text, arrays) cache of cigar_arrays()
        self.rnext = -1
        self.pnext = -1
        self.tlen = 0
        self.seq = ""
        self.qual = None
        self.tags = []            # raw "TAG:TYPE:VALUE" strings, passed through

    # ---- cigar: text is parsed only when somebody looks at it ----
    @property
    def cigar(self):
        if self._cigar is None and self._cigar_s is not None:
            self._cigar = parse_cigar(self._cigar_s)
            self._cigar_s = None
        return self._cigar

    @cigar.setter
    def cigar(self, value):
        self._cigar = tuple(value) if value is not None else None
        self._cigar_s = None

    def cigar_arrays(self):
        """(op codes uint8, lengths int64); straight from the text when the tuple form was never needed."""
        if self._cigar is None and self._cigar_s is not None:
            if self._cigar_a is None or self._cigar_a[0] is not self._cigar_s:
                self._cigar_a = (self._cigar_s, parse_cigar_arrays(self._cigar_s))     # parsed once per text
            return self._cigar_a[1]
        c = self._cigar or ()
        a = np.asarray(c, dtype=np.int64).reshape(-1, 2)
        return a[:, 0].astype(np.uint8), a[:, 1]

    def set_cigar_arrays(self, codes, lens):
        """Assigns the cigar from arrays; the tuple form is only built if somebody asks for it."""
        self._cigar = None
        self._cigar_s = format_cigar_arrays(codes, lens)

    def cigar_text(self):
        return self._cigar_s if self._cigar is None and self._cigar_s is not None else format_cigar(self._cigar)

    # ---- flags ----
    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @is_reverse.setter
    def is_reverse(self, v):
        self.flag = (self.flag | 0x10) if v else (self.flag & ~0x10)

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    # ---- derived coordinates ----
    @property
    def cigarstring(self):
        return format_cigar(self.cigar) if self.cigar else None

    @property
    def qstart(self):
        """Offset of the first aligned base in seq (leading soft clip; hard clips are not in seq)."""
        n = 0
        for op, ln in self.cigar or ():
            if op == BAM_CSOFT_CLIP:
                n += ln
            elif op == BAM_CHARD_CLIP:
                continue
            else:
                break
        return n

    @property
    def qend(self):
        n = len(self.seq or "")
        for op, ln in reversed(self.cigar or ()):
            if op == BAM_CSOFT_CLIP:
                n -= ln
            elif op == BAM_CHARD_CLIP:
                continue
            else:
                break
        return n

    @property
    def query(self):
        return (self.seq or "")[self.qstart:self.qend]

    @property
    def alen(self):
        return sum(ln for op, ln in self.cigar or () if op in _REF_CONSUMING)

    @property
    def aend(self):
        if self.rname < 0 or not self.cigar:
            return None
        return self.pos + self.alen

    @property
    def aligned_pairs(self):
        """[(position in query, reference position)], None on the gapped side; query-relative like pysam 0.7
        (the reference adds the soft-clip offset itself, utils.py:155-166)."""
        out, q, r = [], 0, self.pos
        for op, ln in self.cigar or ():
            if op in (BAM_CMATCH, BAM_CEQUAL, BAM_CDIFF):
                out.extend(zip(range(q, q + ln), range(r, r + ln)))
                q += ln
                r += ln
            elif op == BAM_CINS:
                out.extend((i, None) for i in range(q, q + ln))
                q += ln
            elif op in (BAM_CDEL, BAM_CREF_SKIP):
                out.extend((None, i) for i in range(r, r + ln))
                r += ln
        return out

    def sort_key(self):
        return (self.rname if self.rname >= 0 else 1 << 30, self.pos, self.qname, self.flag)

    def __lt__(self, o):
        return self.sort_key() < o.sort_key()


class Samfile:
    """Text SAM file.  mode "r", or "w"/"wh" with template= (an open Samfile) or header_lines=."""

    def __init__(self, path, mode="r", template=None, header_lines=None):
        self.path, self.mode = path, mode
        self.header_lines = []
        self.references, self.lengths = [], []
        self._tid = {}
        if mode == "r":
            self._fh = open(path, "r")
            self._pending = None
            while True:
                line = self._fh.readline()
                if line.startswith("@"):
                    self._add_header(line.rstrip("\r\n"))
                    continue
                self._pending = line
                break
        elif mode in ("w", "wh"):
            src = template.header_lines if template is not None else (header_lines or [])
            for ln in src:
                self._add_header(ln)
            self._fh = open(path, "w")
            for ln in self.header_lines:
                self._fh.write(ln + "\n")
        else:
            raise ValueError("unsupported mode %r (text SAM only)" % mode)

    def _add_header(self, line):
        self.header_lines.append(line)
        if line.startswith("@SQ"):
            name, ln = None, 0
            for f in line.split("\t")[1:]:
                if f.startswith("SN:"):
                    name = f[3:]
                elif f.startswith("LN:"):
                    ln = int(f[3:])
            if name is not None and name not in self._tid:
                self._tid[name] = len(self.references)
                self.references.append(name)
                self.lengths.append(ln)

    def getrname(self, tid):
        return self.references[tid]

    def gettid(self, name):
        return self._tid.get(name, -1)

    def _tid_of(self, name):
        if name == "*":
            return -1
        if name not in self._tid:            # header-less SAM: register on the fly
            self._tid[name] = len(self.references)
            self.references.append(name)
            self.lengths.append(0)
        return self._tid[name]

    def __iter__(self):
        return self

    def __next__(self):
        while True:
            line = self._pending if self._pending is not None else self._fh.readline()
            self._pending = None
            if line == "":
                raise StopIteration
            line = line.rstrip("\r\n")
            if line == "" or line.startswith("@"):
                continue
            return self._parse(line)

    def _parse(self, line):
        f = line.split("\t")
        if len(f) < 11:
            raise ValueError("SAM record with %d fields: %s" % (len(f), line[:60]))
        a = AlignedRead()
        a.qname, a.flag = f[0], int(f[1])
        a.rname = self._tid_of(f[2])
        a.pos, a.mapq = int(f[3]) - 1, int(f[4])
        a._cigar_s = None if f[5] in ("*", "") else f[5]
        a.rnext = a.rname if f[6] == "=" else self._tid_of(f[6])
        a.pnext, a.tlen = int(f[7]) - 1, int(f[8])
        a.seq = "" if f[9] == "*" else f[9]
        a.qual = None if f[10] == "*" else f[10]
        a.tags = f[11:]
        return a

    def write(self, a):
        rn = self.references[a.rname] if a.rname >= 0 else "*"
        if a.rnext < 0:
            rnext = "*"
        elif a.rnext == a.rname:
            rnext = "="
        else:
            rnext = self.references[a.rnext]
        fields = [a.qname, str(a.flag), rn, str(a.pos + 1), str(a.mapq), a.cigar_text(), rnext,
                  str(a.pnext + 1), str(a.tlen), a.seq if a.seq else "*", a.qual if a.qual else "*"]
        fields.extend(a.tags)
        self._fh.write("\t".join(fields) + "\n")

    def close(self):
        if self._fh is not None:
            self._fh.close()
            self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
