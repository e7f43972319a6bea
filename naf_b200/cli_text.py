"""Text the command-line tools print around the hot path (pure host formatting)."""
from __future__ import annotations

_TYPE_NAMES = {"dna": "DNA", "rna": "RNA", "protein": "protein", "text": "text", 0: "DNA", 1: "RNA", 2: "protein", 3: "text"}


def unexpected_report(info, seq_type="dna") -> bytes:
    """ennaf's stderr report of unexpected input characters (ennaf/src/process.c:75-96), byte for byte."""
    names = ["id", "comment", _TYPE_NAMES[seq_type], "quality"]
    out = []
    for k in range(4):
        counts = list(info.unexpected[k])
        total = sum(counts)
        if not total:
            continue
        out.append(f"input has {total} unexpected {names[k]} characters:\n")
        for c in range(256):
            if counts[c]:
                out.append(f"    '{chr(c)}': {counts[c]}\n" if 32 <= c < 127 else f"    '\\x{c:02X}': {counts[c]}\n")
        if counts[256]:
            out.append(f"    EOF: {counts[256]}\n")
    return "".join(out).encode("latin-1")
