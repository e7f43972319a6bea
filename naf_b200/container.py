"""Host-side .naf container helpers (header, VLE numbers) and the text formatting of the unnaf views
whose payload is tiny.  Mirrors unnaf/src/input.c:31 read_header, unnaf/src/utils.c:117 read_number,
ennaf/src/encoders.c:175 write_variable_length_encoded_number and the print_* functions of
unnaf/src/output.c that only touch the header or a small stream.  Pure host code: nothing here is
on the hot path.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Tuple

SEQ_TYPE_NAMES = ["DNA", "RNA", "protein", "text"]
SECTION_NAMES = ["IDs", "Names", "Lengths", "Mask", "Data", "Quality"]


class NafFormatError(ValueError):
    pass


def put_vle(v: int) -> bytes:
    out = [v & 127]
    v >>= 7
    while v:
        out.append(128 | (v & 127))
        v >>= 7
    return bytes(reversed(out))


def get_vle(buf: bytes, pos: int) -> Tuple[int, int]:
    if pos >= len(buf):
        raise NafFormatError("incomplete or truncated input\n")
    c = buf[pos]; pos += 1
    if c == 128:
        raise NafFormatError("invalid input: error parsing variable length encoded number\n")
    a = 0
    while c & 128:
        if a & (127 << 57):
            raise NafFormatError("invalid input: overflow reading a variable length encoded number\n")
        a = (a << 7) | (c & 127)
        if pos >= len(buf):
            raise NafFormatError("incomplete or truncated input\n")
        c = buf[pos]; pos += 1
    if a & (127 << 57):
        raise NafFormatError("invalid input: overflow reading a variable length encoded number\n")
    return (a << 7) | c, pos


@dataclass
class Header:
    version: int = 1
    seq_type: int = 0
    flags: int = 0
    sep: int = 32
    line_length: int = 0
    n_sequences: int = 0
    title: bytes = b""
    sections: List[Tuple[int, int, int]] = field(default_factory=list)   # (orig, comp, offset) or None per slot

    def has(self, bit: int) -> bool:
        return bool((self.flags >> bit) & 1)


def read_header(buf: bytes) -> Header:
    if len(buf) == 0:
        raise NafFormatError("empty input")
    if len(buf) < 3:
        raise NafFormatError("incomplete or truncated input\n")
    if buf[:3] != b"\x01\xf9\xec":
        raise NafFormatError("not a NAF format\n")
    h = Header()
    pos = 3
    try:
        h.version = buf[pos]; pos += 1
        if h.version < 1 or h.version > 2:
            raise NafFormatError(f"unknown version ({h.version}) of NAF format\n")
        if h.version > 1:
            t = buf[pos]; pos += 1
            if t < 1 or t > 3:
                raise NafFormatError(f"unknown sequence type ({t}) found in NAF file\n")
            h.seq_type = t
        h.flags = buf[pos]; pos += 1
        h.sep = buf[pos]; pos += 1
    except IndexError:
        raise NafFormatError("incomplete or truncated input\n")
    if h.sep < 0x20 or h.sep > 0x7E:
        raise NafFormatError("unsupported name separator character\n")
    h.line_length, pos = get_vle(buf, pos)
    h.n_sequences, pos = get_vle(buf, pos)
    if h.has(6):
        tl, pos = get_vle(buf, pos)
        h.title = bytes(buf[pos:pos + tl]); pos += tl
    h.sections = [None] * 6
    for k, bit in enumerate([5, 4, 3, 2, 1, 0]):
        if not h.has(bit):
            continue
        if h.n_sequences == 0 and pos >= len(buf):
            break
        orig, pos = get_vle(buf, pos)
        comp, pos = get_vle(buf, pos)
        if comp > len(buf) - pos:
            raise NafFormatError("incomplete or truncated input\n")
        h.sections[k] = (orig, comp, pos)
        pos += comp
    return h


def host_view(naf: bytes, view: str):
    """Views answered from the header alone (unnaf.c:395-408, output.c:7-92).  Returns None if the view needs streams."""
    if view not in ("format", "part-list", "sizes", "number", "title", "total-length"):
        return None
    h = read_header(naf)
    if view == "format":
        return f"{SEQ_TYPE_NAMES[h.seq_type]} sequences{' with qualities' if h.has(0) else ''} in NAF format version {h.version}\n".encode()
    if view == "part-list":
        names = ["Title", "IDs", "Names", "Lengths", "Mask", "Data", "Quality"]
        present = [h.has(6), h.has(5), h.has(4), h.has(3), h.has(2), h.has(1), h.has(0)]
        return (", ".join(n for n, p in zip(names, present) if p) + "\n").encode()
    if view == "number":
        return f"{h.n_sequences}\n".encode()
    if view == "sizes":
        out = ""
        if h.has(6):
            out += f"Title: {len(h.title)}\n"
        for k, s in enumerate(h.sections):
            if s is not None:
                ratio = (s[1] / s[0] * 100) if s[0] else float("nan")
                out += f"{SECTION_NAMES[k]}: {s[1]} / {s[0]} ({ratio:.3f}%)\n"
        return out.encode()
    if view == "title":
        return (h.title if h.has(6) else b"") + b"\n"
    if view == "total-length":
        if h.n_sequences == 0 or not h.has(3) or h.sections[4] is None:
            return b""
        return f"{h.sections[4][0]}\n".encode()
    return None


def format_lengths(raw: bytes) -> bytes:
    """unnaf --lengths (output.c:180): merge 0xFFFFFFFF continuation units, one decimal per line."""
    units = struct.unpack(f"<{len(raw) // 4}I", raw)
    out, i, n = [], 0, len(units)
    while i < n:
        v = 0
        while i < n and units[i] == 0xFFFFFFFF:
            v += 0xFFFFFFFF; i += 1
        if i < n:
            v += units[i]
        out.append(str(v)); i += 1
    return ("\n".join(out) + "\n").encode() if out else b""


def format_mask(raw: bytes) -> bytes:
    """unnaf --mask (output.c:222)."""
    out, i, n = [], 0, len(raw)
    while i < n:
        v = 0
        while i < n and raw[i] == 255:
            v += 255; i += 1
        if i < n:
            v += raw[i]
        out.append(str(v)); i += 1
    return ("\n".join(out) + "\n").encode() if out else b""


def format_charcount(raw: bytes) -> bytes:
    """unnaf --charcount (output.c:596-598) from 256 little-endian u64 counts."""
    counts = struct.unpack("<256Q", raw) if raw else [0] * 256
    out = []
    for c in range(256):
        if counts[c]:
            out.append((f"{chr(c)}\t{counts[c]}\n" if 33 <= c < 127 else f"\\x{c:02X}\t{counts[c]}\n"))
    return "".join(out).encode("latin-1")
