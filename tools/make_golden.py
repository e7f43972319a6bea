#!/usr/bin/env python3
"""Regenerates tests/golden/ from the UNMODIFIED reference (needs /root/reference and oracle/_ref built:
`make -C oracle ref`).  Run in the build container only; the fixtures it writes are committed so that
the GPU box (which has no /root/reference) can run the parity tests.

  tests/golden/ref_suite/   the reference's own perl test-suite cases (tests/{alphabet,charcount,small,large}):
                            input file, expected stdout / stderr files copied from the *-ref goldens, plus
                            the .naf the reference ennaf makes for that command line
  tests/golden/cases/       extra inputs (FASTQ, RNA, protein, long mask runs, multi-block streams, levels
                            1/3/19, --long) with the reference-made .naf and sha256 of every unnaf view
  tests/golden/zstd/        zstd frames (reference libzstd at several levels, zstd's decodecorpus) + originals
"""
import ctypes
import gzip
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden")
TMP = "/tmp/make_golden"

sys.path.insert(0, ROOT)
from naf_b200 import synth  # noqa: E402


def run(cmd, stdin=None):
    p = subprocess.run(cmd, input=stdin, capture_output=True, env=dict(os.environ, TMPDIR=TMP))
    return p.returncode, p.stdout, p.stderr


def sha(b):
    return hashlib.sha256(b).hexdigest()


def ref_suite():
    dst = os.path.join(OUT, "ref_suite")
    os.makedirs(dst, exist_ok=True)
    cases = []
    for group_dir in ["alphabet", "charcount", "small", "large"]:
        d = os.path.join(REF, "tests", group_dir)
        os.makedirs(os.path.join(dst, group_dir), exist_ok=True)
        for t in sorted(os.listdir(d)):
            if not t.endswith(".test"):
                continue
            name = t[:-5]
            cmd = open(os.path.join(d, t)).read().strip()
            m = re.match(r"ennaf (.*?)\{GROUP\}\.fa 2>\{TEST\}\.e\.err \| unnaf(.*?)>\{TEST\}\.out 2>\{TEST\}\.u\.err$", cmd)
            assert m, cmd
            eargs, uargs = m.group(1).split(), m.group(2).split()
            group = name.split("-")[0]
            src = os.path.join(d, group + ".fa")
            shutil.copy(src, os.path.join(dst, group_dir, group + ".fa"))
            for ext in ["out", "e.err", "u.err"]:
                shutil.copy(os.path.join(d, f"{name}.{ext}-ref"), os.path.join(dst, group_dir, f"{name}.{ext}"))
            rc, naf, err = run([os.path.join(BIN, "ennaf"), "--binary-stderr", *eargs, src, "-c"])
            assert rc == 0, (cmd, err)
            assert err == open(os.path.join(d, f"{name}.e.err-ref"), "rb").read(), cmd
            rc, out, uerr = run([os.path.join(BIN, "unnaf"), "--binary-stderr", "--binary-stdout", *uargs], naf)
            assert out == open(os.path.join(d, f"{name}.out-ref"), "rb").read(), cmd   # oracle/_ref == the reference's goldens
            open(os.path.join(dst, group_dir, f"{name}.naf"), "wb").write(naf)
            cases.append({"set": group_dir, "name": name, "input": f"{group_dir}/{group}.fa", "ennaf_args": eargs,
                          "unnaf_args": uargs, "naf": f"{group_dir}/{name}.naf"})
    json.dump(cases, open(os.path.join(dst, "manifest.json"), "w"), indent=1)
    print("ref_suite:", len(cases), "cases")


VIEWS = ["fasta", "fastq", "seq", "sequences", "4bit", "ids", "names", "lengths", "mask", "charcount", "total-length",
         "total-mask-length", "number", "format", "part-list", "title"]


def extra_cases():
    dst = os.path.join(OUT, "cases")
    os.makedirs(dst, exist_ok=True)
    rng = np.random.default_rng(7)

    def fasta(records, width):
        out = bytearray()
        for name, seq in records:
            out += b">" + name + b"\n"
            for i in range(0, len(seq), width):
                out += seq[i:i + width] + b"\n"
        return bytes(out)

    def dna(n, alphabet=b"ACGT"):
        return bytes(rng.choice(np.frombuffer(alphabet, dtype=np.uint8), n))

    def soft(seq, runs):
        s = bytearray(seq)
        for a, b in runs:
            s[a:b] = bytes(s[a:b]).lower()
        return bytes(s)

    cases = {}
    # FASTQ with lowercase + IUPAC (mask is stored but never applied on FASTQ output)
    cases["fq_small"] = (synth.fastq(200, 75, seed=3, lowercase=True, iupac=True), ["--fastq"], {})
    cases["fq_c2_shape"] = (synth.fastq(3000, 150, seed=4), [], {})
    # odd total length, mask runs of exactly 255*k and >= 255, leading masked run
    s1 = soft(dna(1021), [(0, 255), (300, 300 + 510), (900, 1021)])
    cases["mask_runs"] = (fasta([(b"m1 leading masked run", s1), (b"m2", soft(dna(777), [(10, 400)]))], 60), [], {})
    cases["odd_total"] = (fasta([(b"a", dna(5)), (b"b", b""), (b"c x y", dna(8))], 4), [], {})
    # RNA / protein / text
    cases["rna"] = (fasta([(b"r1", dna(500, b"ACGU")), (b"r2 rna", soft(dna(333, b"ACGUN"), [(5, 99)]))], 70), ["--rna"], {})
    cases["protein"] = (synth.protein_fasta(300, 120, seed=5), ["--protein"], {})
    cases["text"] = (fasta([(b"t1 some text", b"Hello,World!" * 20), (b"t2", b"lower_and_UPPER>" * 9)], 50), ["--text"], {})
    # multi-block streams (> 128 KB per stream) at several levels; planted repeats so that matches exist
    big = synth.fasta_softmasked(700_000, width=80, seed=6, n_records=3, repeats=True)
    for lvl in ["-1", "-3", "-19"]:
        cases[f"multiblock_l{lvl[1:]}"] = (big, [lvl], {})
    cases["long31"] = (synth.fasta_softmasked(50_000, width=60, seed=8, n_records=2), ["-22", "--long", "31"], {})
    cases["title_linelen"] = (fasta([(b"x", dna(100))], 30), ["--title", "My title", "--line-length", "17"], {})
    cases["ont_iupac"] = (synth.ont_fasta(12, 2000, 9000, seed=9), [], {})
    # CR LF line ends ('\r' is an end-of-line byte to process.c, runs of them collapse): headers with and without comments, an
    # empty record, a blank line, soft-masked runs across line ends, last line without a line end.  (CRs that are NOT line ends: the differential fuzz in tests/test_oracle.py)
    crlf = fasta([(b"c1 comment here", soft(dna(333), [(50, 130)])), (b"c2", b""), (b"c3", dna(61)), (b"c4 x", dna(120))], 60)
    cases["crlf_fasta"] = (crlf.replace(b"\n", b"\r\n")[:-2].replace(b">c3\r\n", b">c3\r\n\r\n"), [], {})
    cases["crlf_protein"] = (synth.protein_fasta(40, 130, seed=11).replace(b"\n", b"\r\n"), ["--protein"], {})

    manifest = []
    for name, (text, eargs, _) in cases.items():
        rc, naf, err = run([os.path.join(BIN, "ennaf"), "--binary-stderr", *eargs, "-c"], text)
        assert rc == 0, (name, err)
        with gzip.GzipFile(os.path.join(dst, name + ".txt.gz"), "wb", mtime=0) as f:
            f.write(text)
        open(os.path.join(dst, name + ".naf"), "wb").write(naf)
        views = {}
        for v in VIEWS:
            for extra in ([], ["--no-mask"], ["--line-length", "33"]):
                rc, out, uerr = run([os.path.join(BIN, "unnaf"), "--binary-stderr", "--binary-stdout", "--" + v, *extra], naf)
                views[" ".join([v] + extra)] = {"rc": rc, "size": len(out), "sha256": sha(out), "stderr": uerr.decode("latin-1")}
        manifest.append({"name": name, "ennaf_args": eargs, "ennaf_stderr": err.decode("latin-1"), "text_size": len(text),
                         "text_sha256": sha(text), "views": views})
    json.dump(manifest, open(os.path.join(dst, "manifest.json"), "w"), indent=1)
    print("cases:", len(manifest))


def zstd_frames():
    dst = os.path.join(OUT, "zstd")
    os.makedirs(dst, exist_ok=True)
    Z = ctypes.CDLL(os.path.join(BIN, "libzstd.so"))
    Z.ZSTD_compressBound.restype = ctypes.c_size_t
    Z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    Z.ZSTD_compress.restype = ctypes.c_size_t
    Z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]

    def comp(b, level):
        cap = Z.ZSTD_compressBound(len(b))
        buf = ctypes.create_string_buffer(cap)
        n = Z.ZSTD_compress(buf, cap, b, len(b), level)
        return buf.raw[:n]

    rng = np.random.default_rng(11)
    data = {
        "empty": b"", "one": b"A", "rle": b"\xff" * 70000,
        "len": np.tile(np.array([150, 0, 0, 0], dtype=np.uint8), 40000).tobytes(),
        "dna4": rng.choice(np.array([0x11, 0x12, 0x14, 0x18, 0x21, 0x22, 0x24, 0x28, 0x41, 0x42, 0x44, 0x48, 0x81, 0x82,
                                     0x84, 0x88], dtype=np.uint8), 150000).tobytes(),
        "qual": (np.clip(np.round(rng.normal(34, 6, 140000)), 2, 40).astype(np.uint8) + 33).tobytes(),
        "ids": b"".join(b"SRR1.%d\0" % i for i in range(15000)),
        "text": open(os.path.join(REF, "zstd/doc/zstd_compression_format.md"), "rb").read(),
    }
    manifest = []
    for name, d in data.items():
        open(os.path.join(dst, name + ".raw"), "wb").write(d)
        for level in [-5, 1, 3, 9, 19]:
            z = comp(d, level)
            fn = f"{name}_l{level}.zst"
            open(os.path.join(dst, fn), "wb").write(z)
            manifest.append({"frame": fn, "raw": name + ".raw", "size": len(d), "sha256": sha(d)})
    # decodecorpus: seeded random frames exercising every block / literal / sequence mode
    dc, dco = os.path.join(TMP, "dc"), os.path.join(TMP, "dco")
    shutil.rmtree(dc, ignore_errors=True); shutil.rmtree(dco, ignore_errors=True)
    os.makedirs(dc); os.makedirs(dco)
    subprocess.run([os.path.join(BIN, "decodecorpus"), "-n400", "-s2024", "-p" + dc, "-o" + dco], capture_output=True, check=True)
    kept = 0
    for f in sorted(os.listdir(dc)):
        z = open(os.path.join(dc, f), "rb").read()
        raw = open(os.path.join(dco, f[:-4]), "rb").read()
        if len(z) + len(raw) > 24000:
            continue
        open(os.path.join(dst, "dc_" + f), "wb").write(z)
        open(os.path.join(dst, "dc_" + f[:-4] + ".raw"), "wb").write(raw)
        manifest.append({"frame": "dc_" + f, "raw": "dc_" + f[:-4] + ".raw", "size": len(raw), "sha256": sha(raw)})
        kept += 1
        if kept >= 120:
            break
    json.dump(manifest, open(os.path.join(dst, "manifest.json"), "w"), indent=1)
    print("zstd frames:", len(manifest))


if __name__ == "__main__":
    os.makedirs(TMP, exist_ok=True)
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    ref_suite()
    extra_cases()
    zstd_frames()
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(OUT) for f in fs)
    print("tests/golden total bytes:", total)
