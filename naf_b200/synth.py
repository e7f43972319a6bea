"""Deterministic synthetic read sets of the shapes BASELINE.json names.

numpy only (PCG64, fixed seeds) so the same bytes come out in the build container and on the GPU
box.  Used by tests (small sizes), tools/make_golden.py and bench.py (full sizes).  Everything is
vectorised: the 10 M-record FASTQ of config 2 (3.3 GB) takes a few tens of seconds.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_IUPAC = np.frombuffer(b"RYSWKMBDHVN", dtype=np.uint8)
_AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


def _qual_table() -> np.ndarray:
    """uniform byte -> 33 + clip(round(N(34, 6)), 2, 40) by inverse CDF."""
    from statistics import NormalDist
    nd = NormalDist(34, 6)
    q = [min(40, max(2, round(nd.inv_cdf((u + 0.5) / 256)))) + 33 for u in range(256)]
    return np.array(q, dtype=np.uint8)


def _digits(i: np.ndarray, nd: int) -> np.ndarray:
    """decimal digits of i (all with exactly nd digits) as an (n, nd) uint8 array"""
    out = np.empty((i.size, nd), dtype=np.uint8)
    v = i.copy()
    for k in range(nd - 1, -1, -1):
        out[:, k] = (v % 10 + 48).astype(np.uint8)
        v //= 10
    return out


def fastq(n_records: int, read_len: int = 150, seed: int = 42, lowercase: bool = False, iupac: bool = False,
          first_index: int = 1) -> bytes:
    """Config 2: `@SRR1.{i} {i}/1`, iid ACGT, Phred-33 quality ~ N(34,6) clipped to [2,40], bare '+' line."""
    return fastq_array(n_records, read_len, seed, lowercase, iupac, first_index).tobytes()


def fastq_array(n_records: int, read_len: int = 150, seed: int = 42, lowercase: bool = False, iupac: bool = False,
                first_index: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    qt = _qual_table()
    parts = []
    lo = first_index
    end = first_index + n_records
    while lo < end:
        nd = len(str(lo))
        hi = min(end, 10 ** nd)
        n = hi - lo
        idx = np.arange(lo, hi, dtype=np.int64)
        dg = _digits(idx, nd)
        head = 6 + nd + 1 + nd + 3            # "@SRR1." digits " " digits "/1\n"
        rec = head + read_len + 3 + read_len + 1
        a = np.empty((n, rec), dtype=np.uint8)
        a[:, 0:6] = np.frombuffer(b"@SRR1.", dtype=np.uint8)
        a[:, 6:6 + nd] = dg
        a[:, 6 + nd] = 32
        a[:, 7 + nd:7 + 2 * nd] = dg
        a[:, 7 + 2 * nd:head] = np.frombuffer(b"/1\n", dtype=np.uint8)
        bases = _ACGT[np.frombuffer(rng.bytes(n * read_len), dtype=np.uint8) & 3].reshape(n, read_len)
        if iupac:
            m = np.frombuffer(rng.bytes(n * read_len), dtype=np.uint8).reshape(n, read_len) < 3
            bases = np.where(m, _IUPAC[np.frombuffer(rng.bytes(n * read_len), dtype=np.uint8) % 11].reshape(n, read_len), bases)
        if lowercase:
            m = np.frombuffer(rng.bytes(n * read_len), dtype=np.uint8).reshape(n, read_len) < 40
            bases = np.where(m, bases | 0x20, bases)
        a[:, head:head + read_len] = bases
        a[:, head + read_len:head + read_len + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
        a[:, head + read_len + 3:rec - 1] = qt[np.frombuffer(rng.bytes(n * read_len), dtype=np.uint8)].reshape(n, read_len)
        a[:, rec - 1] = 10
        parts.append(a.reshape(-1))
        lo = hi
    return np.concatenate(parts) if len(parts) > 1 else parts[0]


def _wrap(seq: np.ndarray, width: int) -> np.ndarray:
    """insert '\\n' after every `width` bases and at the end (if any bases)"""
    n = seq.size
    if n == 0:
        return seq
    nl = (n + width - 1) // width
    out = np.full(n + nl, 10, dtype=np.uint8)
    pos = np.arange(n, dtype=np.int64)
    out[pos + pos // width] = seq
    return out


def _softmask(seq: np.ndarray, rng, run_lo: int, run_hi: int, gap_lo: int, gap_hi: int) -> None:
    n = seq.size
    est = max(4, int(n / ((run_lo + run_hi + gap_lo + gap_hi) / 2) * 1.3) + 8)
    gaps = rng.integers(gap_lo, gap_hi + 1, est)
    runs = rng.integers(run_lo, run_hi + 1, est)
    starts = np.cumsum(gaps + np.concatenate(([0], runs[:-1])))
    ends = starts + runs
    keep = starts < n
    starts, ends = starts[keep], np.minimum(ends[keep], n)
    delta = np.zeros(n + 1, dtype=np.int32)
    np.add.at(delta, starts, 1)
    np.add.at(delta, ends, -1)
    m = np.cumsum(delta[:-1]) > 0
    seq[m] |= 0x20


def fasta_softmasked(n_bases: int, width: int = 60, seed: int = 42, n_records: int = 1, repeats: bool = False,
                     n_gaps: int = 0) -> bytes:
    """Config 5: human-like soft-masked FASTA, ~50 % lowercase in runs U[100,5000], optional 50 kbp N gaps."""
    rng = np.random.default_rng(seed)
    seq = _ACGT[np.frombuffer(rng.bytes(n_bases), dtype=np.uint8) & 3].copy()
    if repeats and n_bases > 40000:
        unit = seq[1000:9000].copy()
        for k in range(6):
            at = int(rng.integers(10000, n_bases - 9000))
            seq[at:at + unit.size] = unit
    for _ in range(n_gaps):
        g = min(50000, n_bases // 20)
        at = int(rng.integers(0, max(1, n_bases - g)))
        seq[at:at + g] = ord("N")
    _softmask(seq, rng, 100, 5000, 100, 5000)
    cuts = np.linspace(0, n_bases, n_records + 1).astype(np.int64)
    parts = []
    for r in range(n_records):
        parts.append(np.frombuffer(b">chr%d synthetic soft-masked\n" % (r + 1), dtype=np.uint8))
        parts.append(_wrap(seq[cuts[r]:cuts[r + 1]], width))
    return np.concatenate(parts).tobytes()


def ont_fasta(n_records: int, len_lo: int = 10000, len_hi: int = 50000, seed: int = 42, width: int = 80) -> bytes:
    """Config 3: ONT-like FASTA, one IUPAC code per ~2000 bases, lowercase runs U[50,2000] every U[200,5000]."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(len_lo, len_hi + 1, n_records)
    total = int(lens.sum())
    seq = _ACGT[np.frombuffer(rng.bytes(total), dtype=np.uint8) & 3].copy()
    k = max(1, total // 2000)
    at = rng.integers(0, total, k)
    seq[at] = _IUPAC[rng.integers(0, 11, k)]
    _softmask(seq, rng, 50, 2000, 200, 5000)
    parts = []
    off = 0
    for i, L in enumerate(lens):
        parts.append(np.frombuffer(b">ont_%d len=%d\n" % (i, L), dtype=np.uint8))
        parts.append(_wrap(seq[off:off + L], width))
        off += int(L)
    return np.concatenate(parts).tobytes()


def protein_fasta(n_records: int, length: int = 300, seed: int = 42, width: int = 60) -> bytes:
    """Config 4: `>sp|P{i:05d}|PROT_{i} some protein`, 20 standard residues iid."""
    rng = np.random.default_rng(seed)
    res = _AA[np.frombuffer(rng.bytes(n_records * length), dtype=np.uint8) % 20].reshape(n_records, length)
    nl = (length + width - 1) // width
    body = np.full((n_records, length + nl), 10, dtype=np.uint8)
    pos = np.arange(length)
    body[:, pos + pos // width] = res
    parts = []
    lo = 0
    while lo < n_records:                       # group by digit count so each group is a rectangular array
        nd = max(5, len(str(lo)))
        nd2 = len(str(lo))
        hi = min(n_records, 10 ** nd2 if lo else 10)
        n = hi - lo
        idx = np.arange(lo, hi, dtype=np.int64)
        d5, d = _digits(idx, nd), _digits(idx, nd2)
        head = 5 + nd + 6 + nd2 + 14
        a = np.empty((n, head + body.shape[1]), dtype=np.uint8)
        a[:, 0:5] = np.frombuffer(b">sp|P", dtype=np.uint8)
        a[:, 5:5 + nd] = d5
        a[:, 5 + nd:11 + nd] = np.frombuffer(b"|PROT_", dtype=np.uint8)
        a[:, 11 + nd:11 + nd + nd2] = d
        a[:, 11 + nd + nd2:head] = np.frombuffer(b" some protein\n", dtype=np.uint8)
        a[:, head:] = body[lo:hi]
        parts.append(a.reshape(-1))
        lo = hi
    return np.concatenate(parts).tobytes()


def fasta_reads(n_records: int = 1000, read_len: int = 150, seed: int = 42) -> bytes:
    """Config 1: `>read{i}`, iid ACGT, one line per record."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    bases = _ACGT[np.frombuffer(rng.bytes(n_records * read_len), dtype=np.uint8) & 3].reshape(n_records, read_len)
    for i in range(n_records):
        out += b">read%d\n" % i + bases[i].tobytes() + b"\n"
    return bytes(out)


def count_bases(kind: str, **kw) -> int:
    if kind == "fastq":
        return kw["n_records"] * kw.get("read_len", 150)
    raise ValueError(kind)
