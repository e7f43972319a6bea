"""Shared test helpers: oracle binding, reference binaries (when present), golden manifests."""
import ctypes as C
import gzip
import hashlib
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref")


class OBuf(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("size", C.c_size_t), ("cap", C.c_size_t)]


class OEncOpts(C.Structure):
    _fields_ = [("seq_type", C.c_int), ("no_mask", C.c_int), ("well_formed", C.c_int), ("strict", C.c_int),
                ("have_line_length", C.c_int), ("line_length", C.c_uint64), ("title", C.c_char_p), ("level", C.c_int),
                ("window_log", C.c_int)]


class ODecOpts(C.Structure):
    _fields_ = [("out_type", C.c_int), ("no_mask", C.c_int), ("have_line_length", C.c_int), ("line_length", C.c_uint64)]


class OStreams(C.Structure):
    _fields_ = [("ids", OBuf), ("comm", OBuf), ("len", OBuf), ("mask", OBuf), ("seq", OBuf), ("qual", OBuf),
                ("n_sequences", C.c_uint64), ("longest_line", C.c_uint64), ("seq_size", C.c_uint64), ("format", C.c_int),
                ("store_mask", C.c_int), ("store_qual", C.c_int), ("unexpected", (C.c_uint64 * 257) * 4)]


OVIEWS = {"default": 0, "format": 1, "part-list": 2, "sizes": 3, "number": 4, "title": 5, "ids": 6, "names": 7, "lengths": 8,
          "total-length": 9, "mask": 10, "total-mask-length": 11, "4bit": 12, "seq": 13, "sequences": 14, "charcount": 15,
          "fasta": 16, "fastq": 17}
SEQ_TYPES = {"dna": 0, "rna": 1, "protein": 2, "text": 3}

_oracle = None


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.ozstd_decompress.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(OBuf), C.c_char_p]
        lib.ozstd_decompress_frame.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(OBuf), C.POINTER(C.c_size_t), C.c_char_p]
        lib.onaf_encode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(OEncOpts), C.POINTER(OBuf), C.POINTER(OBuf), C.c_char_p]
        lib.onaf_decode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(ODecOpts), C.POINTER(OBuf), C.c_char_p]
        lib.onaf_split.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(OEncOpts), C.POINTER(OStreams), C.c_char_p]
        lib.onaf_streams_init.argtypes = [C.POINTER(OStreams)]
        lib.onaf_streams_free.argtypes = [C.POINTER(OStreams)]
        lib.obuf_free.argtypes = [C.POINTER(OBuf)]
        for name in ("onaf_pack4", "onaf_mask_rle"):
            getattr(lib, name).argtypes = [C.c_char_p, C.c_size_t, C.POINTER(OBuf)]
        lib.onaf_unpack4.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(OBuf)]
        lib.onaf_put_vle.argtypes = [C.POINTER(OBuf), C.c_uint64]
        lib.onaf_get_vle.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.c_char_p]

    @staticmethod
    def _take(lib, ob):
        out = C.string_at(ob.data, ob.size) if ob.size else b""
        lib.obuf_free(C.byref(ob))
        return out

    def zstd_decompress(self, z, one_frame=False):
        ob, err = OBuf(), C.create_string_buffer(256)
        if one_frame:
            used = C.c_size_t()
            rc = self.lib.ozstd_decompress_frame(z, len(z), C.byref(ob), C.byref(used), err)
        else:
            rc = self.lib.ozstd_decompress(z, len(z), C.byref(ob), err)
        out = self._take(self.lib, ob)
        if rc:
            raise ValueError(err.value.decode())
        return out

    @staticmethod
    def enc_opts(seq_type="dna", no_mask=False, well_formed=False, strict=False, line_length=None, title=None, window_log=0, **_):
        o = OEncOpts()
        o.seq_type = SEQ_TYPES[seq_type]
        o.no_mask, o.well_formed, o.strict = int(no_mask), int(well_formed), int(strict)
        o.have_line_length, o.line_length = int(line_length is not None), int(line_length or 0)
        o.title = title.encode() if title else None
        o.window_log = window_log
        return o

    def encode(self, text, **kw):
        """-> (naf bytes, stderr report bytes); raises ValueError(die message)"""
        naf, rep, err = OBuf(), OBuf(), C.create_string_buffer(256)
        o = self.enc_opts(**kw)
        rc = self.lib.onaf_encode(text, len(text), C.byref(o), C.byref(naf), C.byref(rep), err)
        a, b = self._take(self.lib, naf), self._take(self.lib, rep)
        if rc:
            raise ValueError(err.value.decode("latin-1"))
        return a, b

    def split(self, text, **kw):
        s, err = OStreams(), C.create_string_buffer(256)
        self.lib.onaf_streams_init(C.byref(s))
        o = self.enc_opts(**kw)
        rc = self.lib.onaf_split(text, len(text), C.byref(o), C.byref(s), err)
        if rc:
            self.lib.onaf_streams_free(C.byref(s))
            raise ValueError(err.value.decode("latin-1"))
        streams = [C.string_at(b.data, b.size) if b.size else b"" for b in (s.ids, s.comm, s.len, s.mask, s.seq, s.qual)]
        info = {"n_sequences": s.n_sequences, "longest_line": s.longest_line, "seq_size": s.seq_size, "format": s.format,
                "store_mask": s.store_mask, "store_qual": s.store_qual,
                "unexpected": [[s.unexpected[k][c] for c in range(257)] for k in range(4)]}
        self.lib.onaf_streams_free(C.byref(s))
        return streams, info

    def decode(self, naf, view="default", no_mask=False, line_length=None):
        out, err = OBuf(), C.create_string_buffer(256)
        o = ODecOpts()
        o.out_type, o.no_mask = OVIEWS[view], int(no_mask)
        o.have_line_length, o.line_length = int(line_length is not None), int(line_length or 0)
        rc = self.lib.onaf_decode(naf, len(naf), C.byref(o), C.byref(out), err)
        a = self._take(self.lib, out)
        if rc:
            raise ValueError(err.value.decode("latin-1"))
        return a


def load_oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
        _oracle = Oracle(C.CDLL(so))
    return _oracle


def have_ref():
    return all(os.access(os.path.join(REF_BIN, b), os.X_OK) for b in ("ennaf", "unnaf"))


def ref_run(tool, args, stdin=b"", tmp="/tmp", timeout=None):
    env = dict(os.environ, TMPDIR=tmp)
    extra = ["--binary-stderr"] + (["--binary-stdout"] if tool == "unnaf" else [])
    p = subprocess.run([os.path.join(REF_BIN, tool), *extra, *args], input=stdin, capture_output=True, env=env, timeout=timeout)
    return p.returncode, p.stdout, p.stderr


def sha(b):
    return hashlib.sha256(b).hexdigest()


def manifest(kind):
    return json.load(open(os.path.join(GOLDEN, kind, "manifest.json")))


def golden(*parts):
    p = os.path.join(GOLDEN, *parts)
    if p.endswith(".gz"):
        return gzip.open(p, "rb").read()
    return open(p, "rb").read()


def parse_ennaf_args(args):
    """reference command-line flags -> keyword arguments shared by the oracle and naf_b200"""
    kw, i = {}, 0
    while i < len(args):
        a = args[i]
        if a in ("--dna", "--rna", "--protein", "--text"):
            kw["seq_type"] = a[2:]
        elif a == "--no-mask":
            kw["no_mask"] = True
        elif a == "--well-formed":
            kw["well_formed"] = True
        elif a == "--strict":
            kw["strict"] = True
        elif a == "--line-length":
            i += 1; kw["line_length"] = int(args[i])
        elif a == "--title":
            i += 1; kw["title"] = args[i]
        elif a == "--long":
            i += 1; kw["window_log"] = int(args[i])
        elif a in ("--fasta", "--fastq"):
            pass
        elif a[0] == "-" and a[1:].isdigit():
            kw["level"] = int(a[1:])
        else:
            raise ValueError(a)
        i += 1
    return kw


def parse_unnaf_args(args):
    kw, i = {"view": "default"}, 0
    while i < len(args):
        a = args[i]
        if a == "--no-mask":
            kw["no_mask"] = True
        elif a == "--line-length":
            i += 1; kw["line_length"] = int(args[i])
        elif a.startswith("--"):
            kw["view"] = a[2:]
        i += 1
    return kw


def claim_huge_ids(naf: bytes) -> bytes:
    """the same .naf with the ids section's original-size field rewritten to 2^40 + 1 (a damaged header)"""
    from naf_b200 import container
    h = container.read_header(naf)
    orig, comp, off = h.sections[0]                      # ids: VLE(orig) VLE(comp) payload at `off`
    head = container.put_vle(orig) + container.put_vle(comp)
    at = off - len(head)
    assert naf[at:off] == head
    return naf[:at] + container.put_vle((1 << 40) + 1) + container.put_vle(comp) + naf[off:]


def replace_section(naf: bytes, k: int, donor: bytes) -> bytes:
    """the same .naf with section k (0 ids .. 5 quality: VLE sizes + payload) taken from `donor`"""
    from naf_b200 import container
    def span(buf):
        orig, comp, off = container.read_header(buf).sections[k]
        head = container.put_vle(orig) + container.put_vle(comp)
        assert buf[off - len(head):off] == head
        return off - len(head), off + comp
    a0, a1 = span(naf)
    d0, d1 = span(donor)
    return naf[:a0] + donor[d0:d1] + naf[a1:]


def claim_records(naf: bytes, n: int) -> bytes:
    """the same .naf with the header's number of sequences rewritten to n"""
    from naf_b200 import container
    h = container.read_header(naf)
    pos = 4 + (1 if h.version > 1 else 0) + 2
    ll, p1 = container.get_vle(naf, pos)
    nn, p2 = container.get_vle(naf, p1)
    assert nn == h.n_sequences
    return naf[:p1] + container.put_vle(n) + naf[p2:]
