"""ctypes binding of include/nafgpu.h.  No torch types cross the boundary: plain pointers and sizes."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

DNA, RNA, PROTEIN, TEXT = 0, 1, 2, 3
(OUT_DEFAULT, OUT_FASTA, OUT_FASTQ, OUT_SEQ, OUT_SEQUENCES, OUT_4BIT, OUT_IDS, OUT_NAMES, OUT_LENGTHS, OUT_MASK,
 OUT_CHARCOUNT) = range(11)

_SEQ_TYPES = {"dna": DNA, "rna": RNA, "protein": PROTEIN, "text": TEXT}
_VIEWS = {"default": OUT_DEFAULT, "fasta": OUT_FASTA, "fastq": OUT_FASTQ, "seq": OUT_SEQ, "sequences": OUT_SEQUENCES,
          "4bit": OUT_4BIT, "ids": OUT_IDS, "names": OUT_NAMES, "lengths": OUT_LENGTHS, "mask": OUT_MASK,
          "charcount": OUT_CHARCOUNT}


class NafGpuError(RuntimeError):
    """Raised with the message the reference would have passed to die() (no prefix)."""

    def __init__(self, code: int, message: str):
        super().__init__(message.rstrip("\n") or f"nafgpu error {code}")
        self.code = code
        self.message = message


class EncOpts(C.Structure):
    _fields_ = [("seq_type", C.c_int32), ("input_format", C.c_int32), ("no_mask", C.c_int32), ("strict", C.c_int32),
                ("well_formed", C.c_int32), ("have_line_length", C.c_int32), ("line_length", C.c_uint64),
                ("level", C.c_int32), ("window_log", C.c_int32), ("title", C.c_char_p), ("general_parser", C.c_int32),
                ("no_block_index", C.c_int32)]


class DecOpts(C.Structure):
    _fields_ = [("out_type", C.c_int32), ("no_mask", C.c_int32), ("have_line_length", C.c_int32),
                ("line_length", C.c_uint64), ("first_record", C.c_uint64), ("n_records", C.c_uint64)]


class EncInfo(C.Structure):
    _fields_ = [("n_sequences", C.c_uint64), ("longest_line", C.c_uint64), ("n_bases", C.c_uint64),
                ("format", C.c_int32), ("reserved", C.c_int32), ("stream_raw", C.c_uint64 * 6),
                ("stream_comp", C.c_uint64 * 6), ("unexpected", (C.c_uint64 * 257) * 4)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernels_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("parser_fallback", C.c_uint32)]


class ShardCounts(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_bases", C.c_uint64), ("longest_line", C.c_uint64), ("n_flips", C.c_uint64),
                ("last_flip", C.c_uint64), ("first_code", C.c_uint8), ("first_case", C.c_uint8), ("last_case", C.c_uint8),
                ("format", C.c_uint8), ("pad", C.c_uint8 * 4)]


class ShardLink(C.Structure):
    _fields_ = [("bases_before", C.c_uint64), ("run_carry", C.c_uint64), ("prev_last_case", C.c_uint8),
                ("next_first_code", C.c_uint8), ("is_last", C.c_uint8), ("store_qual", C.c_uint8), ("pad", C.c_uint8 * 4)]


WRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)

# every symbol include/nafgpu.h declares (tests check the .so exports exactly these)
EXPORTS = [
    "nafgpu_create", "nafgpu_destroy", "nafgpu_last_error", "nafgpu_version", "nafgpu_get_timing", "nafgpu_stream",
    "nafgpu_host_alloc", "nafgpu_host_free", "nafgpu_encode", "nafgpu_decode", "nafgpu_encode_device",
    "nafgpu_decode_device", "nafgpu_zstd_decompress", "nafgpu_zstd_compress", "nafgpu_zstd_compress_level", "nafgpu_split", "nafgpu_profile",
    "nafgpu_profile_report", "nafgpu_shard_begin", "nafgpu_shard_finish", "nafgpu_shard_fetch",
    "nafgpu_record_cuts", "nafgpu_encode_begin", "nafgpu_encode_buffer", "nafgpu_encode_feed", "nafgpu_encode_end", "nafgpu_encode_end_to", "nafgpu_decode_to",
]

_lib = None


def library_path() -> str:
    """NAFGPU_LIB names another build of the same library (A/B measurements of kernel variants); there is still no fallback."""
    return os.environ.get("NAFGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnafgpu.so")


def load_library():
    """Load libnafgpu.so from the package directory (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise NafGpuError(-1, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
    lib = C.CDLL(path)
    u8p, vp, sz = C.POINTER(C.c_uint8), C.c_void_p, C.c_size_t
    lib.nafgpu_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.nafgpu_destroy.argtypes = [vp]
    lib.nafgpu_destroy.restype = None
    lib.nafgpu_last_error.argtypes = [vp]
    lib.nafgpu_last_error.restype = C.c_char_p
    lib.nafgpu_version.restype = C.c_char_p
    lib.nafgpu_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.nafgpu_stream.argtypes = [vp]
    lib.nafgpu_stream.restype = vp
    lib.nafgpu_profile.argtypes = [vp, C.c_int]
    lib.nafgpu_profile_report.argtypes = [vp]
    lib.nafgpu_profile_report.restype = C.c_char_p
    lib.nafgpu_host_alloc.argtypes = [sz, C.POINTER(vp)]
    lib.nafgpu_host_free.argtypes = [vp]
    lib.nafgpu_host_free.restype = None
    lib.nafgpu_encode.argtypes = [vp, vp, sz, C.POINTER(EncOpts), C.POINTER(vp), C.POINTER(sz), C.POINTER(EncInfo)]
    lib.nafgpu_decode.argtypes = [vp, vp, sz, C.POINTER(DecOpts), C.POINTER(vp), C.POINTER(sz)]
    lib.nafgpu_encode_device.argtypes = [vp, vp, sz, C.POINTER(EncOpts), C.POINTER(vp), C.POINTER(sz), C.POINTER(EncInfo)]
    lib.nafgpu_decode_device.argtypes = [vp, vp, sz, vp, C.POINTER(DecOpts), C.POINTER(vp), C.POINTER(sz)]
    lib.nafgpu_zstd_decompress.argtypes = [vp, vp, sz, sz, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    lib.nafgpu_zstd_compress.argtypes = [vp, vp, sz, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    lib.nafgpu_zstd_compress_level.argtypes = [vp, vp, sz, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    lib.nafgpu_split.argtypes = [vp, vp, sz, C.POINTER(EncOpts), C.POINTER(vp * 6), C.POINTER(sz * 6), C.POINTER(EncInfo)]
    lib.nafgpu_shard_begin.argtypes = [vp, vp, sz, C.c_int, C.POINTER(EncOpts), C.POINTER(ShardCounts), C.POINTER(EncInfo)]
    lib.nafgpu_shard_finish.argtypes = [vp, C.POINTER(ShardLink), C.POINTER(C.c_uint64 * 6), C.POINTER(C.c_uint64 * 6)]
    lib.nafgpu_shard_fetch.argtypes = [vp, C.c_int, vp]
    lib.nafgpu_record_cuts.argtypes = [vp, vp, sz, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.nafgpu_encode_begin.argtypes = [vp, C.POINTER(EncOpts), sz]
    lib.nafgpu_encode_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    lib.nafgpu_encode_feed.argtypes = [vp, sz]
    lib.nafgpu_encode_end.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(EncInfo)]
    lib.nafgpu_encode_end_to.argtypes = [vp, WRITE_FN, vp, C.POINTER(sz), C.POINTER(EncInfo)]
    lib.nafgpu_decode_to.argtypes = [vp, vp, sz, C.POINTER(DecOpts), WRITE_FN, vp, C.POINTER(sz)]
    _lib = lib
    return lib


def _bytes_at(addr: int, size: int) -> bytes:
    """copy `size` bytes at `addr` into a bytes object (ctypes.string_at takes a C int: not for texts beyond 2 GB)"""
    if not size:
        return b""
    if size < (1 << 31):
        return C.string_at(addr, size)
    return memoryview((C.c_ubyte * size).from_address(addr)).tobytes()


def _as_ptr(buf):
    """(address, length, keepalive) for bytes / bytearray / memoryview / numpy / torch CPU tensors / int address."""
    if isinstance(buf, tuple):          # (address, nbytes): caller-managed memory (pinned or device)
        return int(buf[0]), int(buf[1]), None
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value or 0, len(buf), buf
    if isinstance(buf, bytearray):
        arr = (C.c_uint8 * len(buf)).from_buffer(buf)
        return C.addressof(arr), len(buf), (arr, buf)
    if hasattr(buf, "data_ptr"):        # torch tensor (CPU, possibly pinned)
        return int(buf.data_ptr()), int(buf.numel() * buf.element_size()), buf
    if hasattr(buf, "ctypes") and hasattr(buf, "nbytes"):   # numpy
        return int(buf.ctypes.data), int(buf.nbytes), buf
    mv = memoryview(buf)
    arr = (C.c_uint8 * mv.nbytes).from_buffer_copy(mv)
    return C.addressof(arr), mv.nbytes, arr


def make_enc_opts(seq_type="dna", fmt=0, no_mask=False, strict=False, well_formed=False, line_length=None, level=1,
                  window_log=0, title: Optional[str] = None, general_parser=False, block_index=True) -> EncOpts:
    o = EncOpts()
    o.seq_type = _SEQ_TYPES.get(seq_type, seq_type) if isinstance(seq_type, str) else int(seq_type)
    o.input_format = {"fasta": 1, "fastq": 2}.get(fmt, fmt) if isinstance(fmt, str) else int(fmt)
    o.no_mask, o.strict, o.well_formed = int(no_mask), int(strict), int(well_formed)
    o.have_line_length = int(line_length is not None)
    o.line_length = int(line_length or 0)
    o.level, o.window_log = int(level), int(window_log)
    o.title = title.encode() if title is not None else None
    o.general_parser = int(general_parser)
    o.no_block_index = int(not block_index)
    return o


def make_dec_opts(view="default", no_mask=False, line_length=None, first_record=0, n_records=0) -> DecOpts:
    o = DecOpts()
    o.first_record, o.n_records = int(first_record), int(n_records)
    o.out_type = _VIEWS[view] if isinstance(view, str) else int(view)
    o.no_mask = int(no_mask)
    o.have_line_length = int(line_length is not None)
    o.line_length = int(line_length or 0)
    return o


class NafGpu:
    """One context = one CUDA stream + device arena + pinned staging on one GPU."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.nafgpu_create(device, C.byref(h))
        if rc != 0:
            raise NafGpuError(rc, (self.lib.nafgpu_last_error(None) or b"").decode("latin-1"))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.nafgpu_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise NafGpuError(rc, (self.lib.nafgpu_last_error(self.h) or b"").decode("latin-1"))

    def timing(self) -> Timing:
        t = Timing()
        self.lib.nafgpu_get_timing(self.h, C.byref(t))
        return t

    def profile(self, enable: bool):
        self.lib.nafgpu_profile(self.h, int(enable))

    def profile_report(self):
        """[(kernel name, launches, total ms)] of the last call (CUDA events on the context's stream)."""
        out = []
        for line in (self.lib.nafgpu_profile_report(self.h) or b"").decode().splitlines():
            name, cnt, ms = line.split("\t")
            out.append((name, int(cnt), float(ms)))
        return out

    # ---- hot path, host buffers
    def encode_raw(self, text, opts: EncOpts):
        """-> (address, size, EncInfo) of the .naf bytes in ctx-owned pinned memory."""
        p, n, keep = _as_ptr(text)
        out, size, info = C.c_void_p(), C.c_size_t(), EncInfo()
        self._check(self.lib.nafgpu_encode(self.h, p, n, C.byref(opts), C.byref(out), C.byref(size), C.byref(info)))
        return out.value or 0, size.value, info

    def encode(self, text, **kw) -> bytes:
        addr, size, _ = self.encode_raw(text, make_enc_opts(**kw))
        return _bytes_at(addr, size)

    def encode_with_info(self, text, **kw):
        addr, size, info = self.encode_raw(text, make_enc_opts(**kw))
        return _bytes_at(addr, size), info

    def decode_raw(self, naf, opts: DecOpts):
        p, n, keep = _as_ptr(naf)
        out, size = C.c_void_p(), C.c_size_t()
        self._check(self.lib.nafgpu_decode(self.h, p, n, C.byref(opts), C.byref(out), C.byref(size)))
        return out.value or 0, size.value

    def decode(self, naf, view="default", no_mask=False, line_length=None, first_record=0, n_records=0) -> bytes:
        addr, size = self.decode_raw(naf, make_dec_opts(view, no_mask, line_length, first_record, n_records))
        return _bytes_at(addr, size)

    def unnaf(self, naf, view="default", no_mask=False, line_length=None) -> bytes:
        """Every unnaf output type, byte-identical to the reference CLI's stdout (unnaf.c:395-447)."""
        from . import container
        naf = bytes(naf) if not isinstance(naf, bytes) else naf
        hv = container.host_view(naf, view)
        if hv is not None:
            return hv
        if view == "total-mask-length":
            h = container.read_header(naf)
            if h.n_sequences == 0:
                return b""
            return f"{sum(self.decode(naf, 'mask')) if h.has(2) else 0}\n".encode()
        raw = self.decode(naf, view, no_mask, line_length)
        if view == "lengths":
            return container.format_lengths(raw)
        if view == "mask":
            return container.format_mask(raw)
        if view == "charcount":
            return container.format_charcount(raw) if raw else b""
        return raw

    # ---- hot path, streamed (what the command-line tools use)
    def encode_pieces(self, pieces, size_hint: int = 0, write=None, **kw) -> bytes:
        """text as an iterable of byte pieces (any sizes) -> .naf; each piece is on its way to the device while the next is read"""
        opts = make_enc_opts(**kw)
        self._check(self.lib.nafgpu_encode_begin(self.h, C.byref(opts), size_hint))
        buf, cap, fill = C.c_void_p(), C.c_size_t(), 0
        self._check(self.lib.nafgpu_encode_buffer(self.h, C.byref(buf), C.byref(cap)))
        for piece in pieces:
            mv, at = memoryview(piece), 0
            while at < len(mv):
                k = min(cap.value - fill, len(mv) - at)
                C.memmove(buf.value + fill, (C.c_char * k).from_buffer_copy(mv[at:at + k]), k)
                fill += k; at += k
                if fill == cap.value:
                    self._check(self.lib.nafgpu_encode_feed(self.h, fill))
                    self._check(self.lib.nafgpu_encode_buffer(self.h, C.byref(buf), C.byref(cap)))
                    fill = 0
        if fill:
            self._check(self.lib.nafgpu_encode_feed(self.h, fill))
        out, size, info = C.c_void_p(), C.c_size_t(), EncInfo()
        if write is not None:                  # the .naf delivered piece by piece as well
            def cb(user, ptr, k):
                write(C.string_at(ptr, k))
                return 0
            fn = WRITE_FN(cb)
            self._check(self.lib.nafgpu_encode_end_to(self.h, fn, None, C.byref(size), C.byref(info)))
            return b""
        self._check(self.lib.nafgpu_encode_end(self.h, C.byref(out), C.byref(size), C.byref(info)))
        return _bytes_at(out.value or 0, size.value)

    def decode_to(self, naf, write, view="default", no_mask=False, line_length=None) -> int:
        """unnaf with the text handed to write(bytes) piece by piece, in order; returns the number of bytes delivered"""
        p, n, keep = _as_ptr(naf)
        opts = make_dec_opts(view, no_mask, line_length)

        def cb(user, ptr, k):
            write(C.string_at(ptr, k))
            return 0
        fn, total = WRITE_FN(cb), C.c_size_t()
        self._check(self.lib.nafgpu_decode_to(self.h, p, n, C.byref(opts), fn, None, C.byref(total)))
        return total.value

    # ---- hot path, device-resident
    def encode_device(self, d_ptr: int, n: int, opts: EncOpts):
        out, size, info = C.c_void_p(), C.c_size_t(), EncInfo()
        self._check(self.lib.nafgpu_encode_device(self.h, d_ptr, n, C.byref(opts), C.byref(out), C.byref(size), C.byref(info)))
        return out.value or 0, size.value, info

    def decode_device(self, d_ptr: int, n: int, host_copy, opts: DecOpts):
        hp = _as_ptr(host_copy)[0] if host_copy is not None else None
        out, size = C.c_void_p(), C.c_size_t()
        self._check(self.lib.nafgpu_decode_device(self.h, d_ptr, n, hp, C.byref(opts), C.byref(out), C.byref(size)))
        return out.value or 0, size.value

    # ---- one file from several shards (driven by naf_b200.sharded)
    def shard_begin(self, text, opts: EncOpts, on_device: bool = False):
        """-> (ShardCounts, EncInfo); text: host buffer, or (device address, nbytes) with on_device=True"""
        p, n, keep = _as_ptr(text)
        counts, info = ShardCounts(), EncInfo()
        self._check(self.lib.nafgpu_shard_begin(self.h, p, n, int(on_device), C.byref(opts), C.byref(counts), C.byref(info)))
        return counts, info

    def shard_finish(self, link: ShardLink):
        """-> (raw sizes[6], body sizes[6]) of this shard's streams"""
        raw, body = (C.c_uint64 * 6)(), (C.c_uint64 * 6)()
        self._check(self.lib.nafgpu_shard_finish(self.h, C.byref(link), C.byref(raw), C.byref(body)))
        return list(raw), list(body)

    def shard_fetch(self, stream: int, dst_address: int):
        self._check(self.lib.nafgpu_shard_fetch(self.h, stream, dst_address))

    def record_cuts(self, text, pieces: int, on_device: bool = False):
        """record-aligned cut points of a text, found on the GPU: [0, c1, ..., n] (pieces + 1 offsets)"""
        p, n, keep = _as_ptr(text)
        cuts = (C.c_uint64 * (pieces + 1))()
        self._check(self.lib.nafgpu_record_cuts(self.h, p, n, int(on_device), pieces, cuts))
        return list(cuts)

    # ---- stages
    def zstd_decompress(self, frame, expected_size: int = 0, one_frame: bool = False) -> bytes:
        p, n, keep = _as_ptr(frame)
        out, size = C.c_void_p(), C.c_size_t()
        self._check(self.lib.nafgpu_zstd_decompress(self.h, p, n, expected_size, int(one_frame), C.byref(out), C.byref(size)))
        return _bytes_at(out.value, size.value)

    def zstd_compress(self, data, window_log: int = 0, level: int = 0) -> bytes:
        p, n, keep = _as_ptr(data)
        out, size = C.c_void_p(), C.c_size_t()
        self._check(self.lib.nafgpu_zstd_compress_level(self.h, p, n, window_log, level, C.byref(out), C.byref(size)))
        return _bytes_at(out.value, size.value)

    def split(self, text, **kw):
        """-> (list of six stream byte strings, EncInfo)"""
        p, n, keep = _as_ptr(text)
        ptrs, sizes, info = (C.c_void_p * 6)(), (C.c_size_t * 6)(), EncInfo()
        opts = make_enc_opts(**kw)
        self._check(self.lib.nafgpu_split(self.h, p, n, C.byref(opts), C.byref(ptrs), C.byref(sizes), C.byref(info)))
        return [_bytes_at(ptrs[k], sizes[k]) for k in range(6)], info


_default: Optional[NafGpu] = None


def _ctx() -> NafGpu:
    global _default
    if _default is None:
        _default = NafGpu()
    return _default


def ennaf(text, **kw) -> bytes:
    """FASTA/FASTQ text -> .naf bytes (ennaf/src/ennaf.c:433 main, minus argv and file handling)."""
    return _ctx().encode(text, **kw)


def unnaf(naf, view="default", no_mask=False, line_length=None) -> bytes:
    """.naf bytes -> text view (unnaf/src/unnaf.c:356 main, minus argv and file handling)."""
    return _ctx().unnaf(naf, view, no_mask, line_length)
