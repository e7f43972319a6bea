"""not-gpu: the per-chunk logic of the canonical-input parser (naf_b200/csrc/naf_fast_hd.cuh) run on the CPU by
tests/emu/emu_parse.cpp, against the oracle's restatement of process.c / encoders.c.

Contract under test: for ANY input the fast parser either (a) declares it non-canonical (the library then redoes
the split with the general FSM parser) or (b) produces exactly the oracle's streams.  Canonical inputs of the
BASELINE shapes must take (b)."""
import os
import random
import struct
import subprocess

import numpy as np
import pytest

import helpers
from naf_b200 import synth

ROOT = helpers.ROOT
EXE = os.path.join(ROOT, "tests", "_build", "emu_parse")
SEQ_TYPES = {"dna": 0, "rna": 1, "protein": 2, "text": 3}


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "emu", "emu_parse.cpp")
    deps = [src, os.path.join(ROOT, "naf_b200/csrc/naf_fast_hd.cuh"), os.path.join(ROOT, "naf_b200/csrc/zstd_hd.cuh")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-Wall", "-Wno-unused-function", "-o", EXE, src], check=True)
    return EXE


def run_emu(emu, tmp_path, text, seq_type="dna", no_mask=False):
    """-> None if the fast parser declines the input, else dict of raw streams"""
    inp, pre = str(tmp_path / "in.txt"), str(tmp_path / "out")
    with open(inp, "wb") as f:
        f.write(text)
    p = subprocess.run([emu, inp, pre, str(SEQ_TYPES[seq_type]), str(int(no_mask))], capture_output=True)
    if p.returncode == 3:
        return None
    assert p.returncode == 0, p.stderr
    out = {k: open(pre + "." + k, "rb").read() for k in ("ids", "comm", "bases", "qual", "recend")}
    n_rec, longest, end_state = (int(x) for x in open(pre + ".info").read().split())
    out.update(n_rec=n_rec, longest=longest, end_state=end_state)
    return out


def check(emu, oracle, tmp_path, text, must_accept=False, **kw):
    got = run_emu(emu, tmp_path, text, kw.get("seq_type", "dna"), kw.get("no_mask", False))
    if got is None:
        assert not must_accept, "canonical input was declined by the fast parser"
        return False
    try:
        want, info = oracle.split(text, **kw)
    except ValueError as e:
        raise AssertionError(f"fast parser accepted input the reference rejects: {e}")
    assert got["ids"] == want[0]
    assert got["comm"] == want[1]
    ends = struct.unpack(f"<{got['n_rec']}Q", got["recend"])
    lens = [e - (ends[i - 1] if i else 0) for i, e in enumerate(ends)]
    assert struct.pack(f"<{len(lens)}I", *lens) == want[2]
    packed = kw.get("seq_type", "dna") in ("dna", "rna")
    import ctypes as C
    if packed:
        ob = helpers.OBuf()
        oracle.lib.onaf_pack4(got["bases"], len(got["bases"]), C.byref(ob))
        assert oracle._take(oracle.lib, ob) == want[4]
        if info["store_mask"]:
            ob = helpers.OBuf()
            oracle.lib.onaf_mask_rle(got["bases"], len(got["bases"]), C.byref(ob))
            assert oracle._take(oracle.lib, ob) == want[3]
    else:
        assert got["bases"] == want[4]
    if info["store_qual"]:
        assert got["qual"] == want[5]
    assert got["n_rec"] == info["n_sequences"]
    assert got["longest"] == info["longest_line"], (got["longest"], info["longest_line"])
    assert all(all(v == 0 for v in row) for row in info["unexpected"])
    return True


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
        assert check(emu, oracle, tmp_path, text, must_accept=True, **kw)
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
        b">x\n" + b"\n".join(seq(1) for _ in range(400)) + b"\n", b">x\n" + seq(70000) + b"\n>y\n" + seq(3) + b"\n",
        b"@r\nACGT\n+\nIIII\n", b"@r\nACGT\n+\nIIII", b"@r c c\nA\n+r c c\nI\n@s\nCC\n+\n@@\n", b"@r\nACGT\n+\n@III\n@s\nAAAA\n+\n+III\n",
    ]
    for w in (1, 2, 59, 60, 61, 127, 128, 16383, 16384, 16385):
        texts.append(b"".join(b">r%d some comment\n" % i + b"\n".join(seq(5 * w)[k:k + w] for k in range(0, 5 * w, w)) + b"\n" for i in range(5)))
    for L in (1, 2, 31, 32, 33, 63, 64, 65, 200, 5000):
        texts.append(b"".join(b"@q%d %d/1\n" % (i, i) + seq(L) + b"\n+\n" + bytes(rng.integers(33, 127, L).astype(np.uint8)) + b"\n" for i in range(300)))
    for t in texts:
        assert check(emu, oracle, tmp_path, t, must_accept=True), t[:60]
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
        b"@r\nAC\n+\nI\xc3\n",
    ]
    for t in muts:
        declined += not check(emu, oracle, tmp_path, t)
    assert declined >= 12
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
