"""not-gpu: the phases of the single-pass encode transform (naf_b200/csrc/naf_fused_hd.cuh) run on the CPU by
tests/emu/emu_fused.cpp, against the oracle's restatement of process.c / encoders.c.

Contract under test: for ANY input the transform either (a) declares it non-canonical (the library then redoes the split
with the general FSM parser) or (b) produces exactly the oracle's streams.  Canonical inputs of the BASELINE shapes must
take (b).  Every case runs with several emulated CTA sizes, so that the partition of chunks / segments / pieces over
threads is exercised."""
import os
import random
import subprocess

import numpy as np
import pytest

import helpers
from naf_b200 import synth

ROOT = helpers.ROOT
EXE = os.path.join(ROOT, "tests", "_build", "emu_fused")
SEQ_TYPES = {"dna": 0, "rna": 1, "protein": 2, "text": 3}


def _build(exe, extra):
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "emu", "emu_fused.cpp")
    deps = [src] + [os.path.join(ROOT, "naf_b200/csrc", f) for f in ("naf_fused_hd.cuh", "naf_fast_hd.cuh", "zstd_hd.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-Wall", "-Wno-unused-function", *extra, "-o", exe, src], check=True)
    return exe


@pytest.fixture(scope="module")
def emu():
    return _build(EXE, [])


@pytest.fixture(scope="module")
def emu_small():
    """the same phases over 512-byte tiles: inputs of a few KB cross dozens of tile boundaries"""
    return _build(EXE + "_t512", ["-DFT_TILE_BYTES=512"])


def mask_units_from_casebits(casebits, n_bases):
    """encoders.c:98-151 + ennaf.c:511 over one case bit per base"""
    bits = np.unpackbits(np.frombuffer(casebits, dtype=np.uint8), bitorder="little")[:n_bases]
    if n_bases == 0:
        return b""
    edges = np.flatnonzero(np.diff(bits)) + 1
    starts = np.concatenate(([0], edges))
    ends = np.concatenate((edges, [n_bases]))
    runs = list((ends - starts).tolist())
    if bits[0] == 1:
        runs.insert(0, 0)
    out = bytearray()
    for L in runs:
        out += b"\xff" * (L // 255) + bytes([L % 255])
    return bytes(out)


def run_emu(emu, tmp_path, text, seq_type="dna", no_mask=False, threads=64):
    """-> None if the transform declines the input, else dict of raw streams"""
    inp, pre = str(tmp_path / "in.txt"), str(tmp_path / "out")
    with open(inp, "wb") as f:
        f.write(text)
    p = subprocess.run([emu, inp, pre, str(SEQ_TYPES[seq_type]), str(int(no_mask)), str(threads)], capture_output=True)
    if p.returncode == 3:
        return None
    assert p.returncode == 0, (p.returncode, p.stderr)
    out = {k: open(pre + "." + k, "rb").read() for k in ("ids", "comm", "seq", "qual", "len", "casebits")}
    n_rec, longest, end_state, n_bases = (int(x) for x in open(pre + ".info").read().split())
    out.update(n_rec=n_rec, longest=longest, end_state=end_state, n_bases=n_bases)
    return out


def check(emu, oracle, tmp_path, text, must_accept=False, threads=(64,), **kw):
    accepted = None
    for nt in threads:
        got = run_emu(emu, tmp_path, text, kw.get("seq_type", "dna"), kw.get("no_mask", False), nt)
        if got is None:
            assert not must_accept, "canonical input was declined by the fused transform"
            assert accepted in (None, False)
            accepted = False
            continue
        assert accepted in (None, True)
        accepted = True
        try:
            want, info = oracle.split(text, **kw)
        except ValueError as e:
            raise AssertionError(f"fused transform accepted input the reference rejects: {e}")
        assert got["ids"] == want[0]
        assert got["comm"] == want[1]
        assert got["len"] == want[2]
        assert got["seq"] == want[4]
        assert got["n_bases"] == info["seq_size"]
        if info["store_mask"]:
            assert mask_units_from_casebits(got["casebits"], got["n_bases"]) == want[3]
        if info["store_qual"]:
            assert got["qual"] == want[5]
        assert got["n_rec"] == info["n_sequences"]
        assert got["longest"] == info["longest_line"], (got["longest"], info["longest_line"])
        assert all(all(v == 0 for v in row) for row in info["unexpected"])
    return accepted


def test_canonical_baseline_shapes(emu, oracle, tmp_path):
    cases = [
        (synth.fasta_reads(300, 150, seed=1), {}),
        (synth.fastq(700, 150, seed=2), {}),
        (synth.fastq(300, 151, seed=3, lowercase=True, iupac=True), {}),
        (synth.fastq(200, 37, seed=4), {"no_mask": True}),
        (synth.ont_fasta(6, 1000, 9000, seed=5), {}),
        (synth.fasta_softmasked(120_000, width=60, seed=6, n_records=3, repeats=True, n_gaps=2), {}),
        (synth.protein_fasta(300, 300, seed=7), {"seq_type": "protein"}),
        (synth.protein_fasta(100, 77, seed=8), {"seq_type": "protein", "no_mask": True}),
        (synth.protein_fasta(100, 77, seed=8), {"seq_type": "text"}),
        (synth.fasta_reads(50, 150, seed=9).replace(b"T", b"U"), {"seq_type": "rna"}),
    ]
    for text, kw in cases:
        assert check(emu, oracle, tmp_path, text, must_accept=True, threads=(64, 512, 8), **kw)
        assert check(emu, oracle, tmp_path, text[:-1], must_accept=True, **kw)        # no final newline


def test_canonical_edge_cases(emu, oracle, tmp_path):
    rng = np.random.default_rng(11)

    def seq(n):
        return bytes(np.frombuffer(b"ACGTacgtNRYKM-", dtype=np.uint8)[rng.integers(0, 14, n)])
    texts = [
        b">a\n", b">a", b">", b">\n", b">a b\n", b">a b", b">a \n", b"> b\nAC\n", b">a\nACGT", b">a\nACGT\n>b\n>c\nAC\n\n\nGT\n",
        b">a\n\n\nAC\n>b\n", b"\n\n>x y z\nACGT\n", b">" + b"n" * 70 + b" " + b"c" * 200 + b"\n" + seq(500) + b"\n",
        b">" + b"n" * 3000 + b"\n" + seq(50) + b"\n>second " + b"k" * 3900 + b"\nAC\n",
        b">x\n" + b"\n".join(seq(63) for _ in range(300)) + b"\n", b">x\n" + b"\n".join(seq(64) for _ in range(300)) + b"\n",
        b">x\n" + seq(70000) + b"\n>y\n" + seq(3) + b"\n", b">x\n" + seq(16381) + b"\n" + seq(5) + b"\n", b">x\n" + seq(16380) + b"\n" + seq(5) + b"\n",
        b"@r\nACGT\n+\nIIII\n", b"@r\nACGT\n+\nIIII", b"@r c c\nA\n+r c c\nI\n@s\nCC\n+\n@@\n", b"@r\nACGT\n+\n@III\n@s\nAAAA\n+\n+III\n",
        b" \n\n" * 9000 + b">late start\nACGT\n",
    ]
    for w in (17, 59, 60, 61, 127, 128, 16383, 16384, 16385):
        texts.append(b"".join(b">r%d some comment\n" % i + b"\n".join(seq(5 * w)[k:k + w] for k in range(0, 5 * w, w)) + b"\n" for i in range(5)))
    for L in (31, 32, 33, 63, 64, 65, 200, 5000, 40000):
        texts.append(b"".join(b"@q%d %d/1\n" % (i, i) + seq(L) + b"\n+\n" + bytes(rng.integers(33, 127, L).astype(np.uint8)) + b"\n" for i in range(300 if L < 1000 else 7)))
    for t in texts:
        assert check(emu, oracle, tmp_path, t, must_accept=True, threads=(64, 16)), t[:60]
    # lines shorter than 16 bytes on average exceed the per-tile tables: declined (general parser), never wrong
    for t in (b">x\n" + b"\n".join(seq(1) for _ in range(4000)) + b"\n", b"".join(b"@q\nA\n+\nI\n" for _ in range(3000)),
              b"".join(b"@q%d %d/1\n" % (i, i) + seq(12) + b"\n+\n" + b"I" * 12 + b"\n" for i in range(3000))):
        check(emu, oracle, tmp_path, t)
    # a header whose first space is further back than the bounded look-back (C6): declined, never wrong
    check(emu, oracle, tmp_path, b">" + b"n" * 20000 + b"\n" + seq(50) + b"\n>second " + b"k" * 17000 + b"\nAC\n")


def test_non_canonical_is_declined_or_exact(emu, oracle, tmp_path):
    """CR/LF, tabs, blank lines in FASTQ, control bytes, unexpected codes, truncated records ...: never a wrong answer"""
    declined = 0
    base_fa, base_fq = synth.fasta_reads(40, 150, seed=21), synth.fastq(40, 150, seed=22)
    muts = [
        base_fa.replace(b"\n", b"\r\n"), base_fq.replace(b"\n", b"\r\n"), base_fa.replace(b"read7", b"read7\tx"),
        base_fa[:3000] + b"Z" + base_fa[3001:], base_fa[:3001] + b" " + base_fa[3001:], base_fq[:5000] + b"\n" + base_fq[5000:],
        base_fq[:-200], base_fq.replace(b"\n+\n", b"\n-\n", 1), base_fa[:2500] + b"\x7f" + base_fa[2501:], base_fa[:2500] + b"\xff" + base_fa[2501:],
        base_fa[:2500] + b"\x00" + base_fa[2501:], base_fq + b"\n", base_fa.replace(b"T", b"U"), b">a\nAC>GT\n", b"@r\nAC\n+\nI\n", b"@r\nAC\n+\nI I\n",
        b"@r\nAC\n+\nI\xc3\n", b"@r\nAC\n+\r\nII\n", b"@r\nACG\n+\nII\n", b"@r\nAC\n+\nIII\n@s\nA\n+\nI\n",
    ]
    for t in muts:
        declined += not check(emu, oracle, tmp_path, t)
    assert declined >= 13
    rng = random.Random(99)
    import test_gpu_encode as tge
    for it in range(400):
        text = tge._fuzz_fastq(rng) if rng.random() < 0.4 else tge._fuzz_fasta(rng)
        kw = {"seq_type": rng.choice(["dna", "rna", "protein", "text"])}
        if rng.random() < 0.25:
            kw["no_mask"] = True
        try:
            oracle.split(text, **kw)
        except ValueError:
            assert run_emu(emu, tmp_path, text, kw["seq_type"], kw.get("no_mask", False)) is None, text
            continue
        check(emu, oracle, tmp_path, text, **kw)


def test_crlf_fasta_is_canonical(emu, emu_small, oracle, tmp_path):
    """CR LF FASTA splits like LF FASTA (process.c treats '\\r' as an end-of-line byte and collapses runs of them): accepted, and
    exact -- also when the CR and the LF fall into different tiles, when the file ends in CR LF or without a line end, with empty
    sequences and blank lines.  A CR anywhere else is declined (or exact), never wrong."""
    rng = np.random.default_rng(11)

    def seq(n):
        return bytes(np.frombuffer(b"ACGTacgtNRYKM-", dtype=np.uint8)[rng.integers(0, 14, n)])
    texts = [synth.fasta_reads(100, 150, seed=1), synth.ont_fasta(4, 1000, 5000, seed=5),
             synth.fasta_softmasked(30_000, width=60, seed=6, n_records=3, repeats=True, n_gaps=2),
             b">a\n>b c\n\n>d\nAC\n\nGT\n>e\n", b">only\n", b">x y z\nACGT"]
    for L in (15, 16, 17, 60, 61, 62, 63, 64, 127, 510, 511, 512):
        texts.append(b"".join(b">q%d c%d\n" % (i, i) + seq(L) + b"\n" + seq(L // 2 + 1) + b"\n" for i in range(60)))
    for t in texts:
        crlf = t.replace(b"\n", b"\r\n")
        for cut in (0, 1, 2):
            u = crlf[:len(crlf) - cut]
            assert check(emu, oracle, tmp_path, u, must_accept=not u.endswith(b"\r"), threads=(64,)) or u.endswith(b"\r"), u[:60]
            check(emu_small, oracle, tmp_path, u, threads=(16,))
    assert check(emu_small, oracle, tmp_path, texts[0].replace(b"\n", b"\r\n"), must_accept=True, threads=(8,))
    for kw, t in (({"seq_type": "protein"}, synth.protein_fasta(60, 300, seed=7)), ({"seq_type": "text", "no_mask": True}, synth.protein_fasta(60, 77, seed=8))):
        assert check(emu_small, oracle, tmp_path, t.replace(b"\n", b"\r\n"), must_accept=True, threads=(16,), **kw)
    # stray CRs: mid-line, doubled, CR without LF as the only line end
    base = synth.fasta_reads(30, 150, seed=3)
    for t in (base.replace(b"\n", b"\r\r\n", 3), base[:1000] + b"\r" + base[1000:], base.replace(b"\n", b"\r"), base.replace(b"\n", b"\n\r", 2)):
        check(emu, oracle, tmp_path, t)
        check(emu_small, oracle, tmp_path, t, threads=(16,))


def test_small_tiles_cross_every_boundary(emu_small, oracle, tmp_path):
    """512-byte tiles: records, lines, headers, 32-base pieces and shared nibbles straddle tile boundaries all the time"""
    rng = np.random.default_rng(5)

    def seq(n):
        return bytes(np.frombuffer(b"ACGTacgtNRYKM-", dtype=np.uint8)[rng.integers(0, 14, n)])
    accepted = 0
    texts = [synth.fastq(120, 150, seed=2), synth.fastq(90, 151, seed=3, lowercase=True, iupac=True), synth.fasta_reads(100, 150, seed=1),
             synth.ont_fasta(4, 1000, 5000, seed=5), synth.fasta_softmasked(30_000, width=60, seed=6, n_records=3, repeats=True, n_gaps=2)]
    for L in (31, 33, 64, 100, 255, 256, 257, 511, 512, 513, 700, 2000):
        texts.append(b"".join(b"@q%d %d/1\n" % (i, i) + seq(L) + b"\n+\n" + bytes(rng.integers(33, 127, L).astype(np.uint8)) + b"\n" for i in range(40)))
        texts.append(b"".join(b">q%d c%d\n" % (i, i) + seq(L) + b"\n" + seq(L // 2 + 1) + b"\n" for i in range(40)))
    for hl in (100, 511, 512, 1500):                      # headers longer than a tile
        texts.append(b"".join(b"@" + b"n" * hl + b" " + b"c" * hl + b"\n" + seq(70) + b"\n+\n" + b"I" * 70 + b"\n" for _ in range(6)))
        texts.append(b"".join(b">" + b"n" * hl + b"\n" + seq(70) + b"\n" for _ in range(6)))
    for t in texts:
        for cut in (0, 1):
            accepted += bool(check(emu_small, oracle, tmp_path, t[:len(t) - cut], threads=(16, 8)))
    assert accepted >= len(texts)
    for kw, t in (({"seq_type": "protein"}, synth.protein_fasta(60, 300, seed=7)), ({"seq_type": "text", "no_mask": True}, synth.protein_fasta(60, 77, seed=8)),
                  ({"seq_type": "rna"}, synth.fasta_reads(50, 150, seed=9).replace(b"T", b"U"))):
        assert check(emu_small, oracle, tmp_path, t, threads=(16,), **kw)
    rng2 = random.Random(7)
    import test_gpu_encode as tge
    for it in range(300):
        text = tge._fuzz_fastq(rng2) if rng2.random() < 0.4 else tge._fuzz_fasta(rng2)
        kw = {"seq_type": rng2.choice(["dna", "rna", "protein", "text"])}
        try:
            oracle.split(text, **kw)
        except ValueError:
            assert run_emu(emu_small, tmp_path, text, kw["seq_type"], False, 16) is None, text
            continue
        check(emu_small, oracle, tmp_path, text, threads=(16,), **kw)


def _canonical_fuzz(rng):
    """random CANONICAL input: LF only, expected codes only, no blank lines in FASTQ; odd shapes otherwise"""
    alpha = b"ACGTacgtNnRYKMSWBDHV-"

    def seq(n):
        return bytes(rng.choice(alpha) for _ in range(n))

    def header():
        name = bytes(rng.randrange(33, 127) for _ in range(rng.choice((0, 1, 3, 10, 40, 300))))
        name = name.replace(b">", b"x")
        r = rng.random()
        if r < 0.3:
            return name
        comm = bytes(rng.randrange(32, 127) for _ in range(rng.choice((0, 1, 5, 30, 600))))
        return name + b" " + comm
    out = bytearray()
    if rng.random() < 0.5:
        for _ in range(rng.randrange(1, 30)):
            L = rng.choice((1, 2, 31, 32, 33, 64, 100, 300, 1000))
            out += b"@" + header() + b"\n" + seq(L) + b"\n+" + (b"" if rng.random() < 0.7 else b"x y") + b"\n" + bytes(rng.randrange(33, 127) for _ in range(L)) + b"\n"
    else:
        for _ in range(rng.randrange(1, 20)):
            out += b">" + header() + b"\n"
            w = rng.choice((1, 7, 60, 61, 200, 1000))
            for _ in range(rng.choice((0, 1, 2, 5, 30))):
                out += seq(rng.choice((0, w, w, w, rng.randrange(1, w + 1)))) + b"\n"
    if rng.random() < 0.3 and out.endswith(b"\n"):
        out = out[:-1]
    return bytes(out)


def test_canonical_fuzz_small_tiles(emu_small, emu, oracle, tmp_path):
    rng = random.Random(1234)
    acc = 0
    for it in range(250):
        text = _canonical_fuzz(rng)
        if not text.strip():
            continue
        try:
            oracle.split(text)
        except ValueError:
            assert run_emu(emu_small, tmp_path, text, "dna", False, 16) is None, text
            continue
        acc += bool(check(emu_small, oracle, tmp_path, text, threads=(16,)))
        assert check(emu, oracle, tmp_path, text, threads=(64,)) is not None
    assert acc > 100
