"""GPU parity, encode side.  Stream level: the six streams the GPU parser/packer produce are byte-identical
to the oracle's restatement of process.c/encoders.c.  File level: the .naf we emit is decoded by the
oracle and by the UNMODIFIED reference unnaf back to the reference's pinned output (north_star: "the
encode path emits a format-valid .naf that the reference unnaf decodes back to the byte-identical input")."""
import os
import random

import numpy as np
import pytest

import helpers
import naf_b200
from naf_b200 import synth

pytestmark = pytest.mark.gpu

STREAMS = ["ids", "comments", "lengths", "mask", "sequence", "quality"]


def check_split(gpu, oracle, text, expect_fast=None, **kw):
    """both parsers: the canonical-input one (with its fallback) and the general FSM one must give the oracle's streams"""
    _check_split(gpu, oracle, text, **kw)
    if expect_fast is not None:
        assert (gpu.timing().parser_fallback == 0) == expect_fast, "fast parser %s this input" % ("declined" if expect_fast else "accepted")
    _check_split(gpu, oracle, text, general_parser=True, **kw)


def _check_split(gpu, oracle, text, general_parser=False, **kw):
    gkw = dict(kw, general_parser=general_parser)
    try:
        want, winfo = oracle.split(text, **kw)
    except ValueError as e:
        with pytest.raises(naf_b200.NafGpuError) as ge:
            gpu.split(text, **gkw)
        assert ge.value.message == str(e), (ge.value.message, str(e))
        return
    got, info = gpu.split(text, **gkw)
    if not winfo["store_mask"]:
        want[3] = b""
    if not winfo["store_qual"]:
        want[5] = b""
    for k in range(6):
        assert got[k] == want[k], (STREAMS[k], kw, len(got[k]), len(want[k]), got[k][:80], want[k][:80])
    assert info.n_sequences == winfo["n_sequences"]
    assert info.longest_line == winfo["longest_line"], (info.longest_line, winfo["longest_line"])
    assert info.n_bases == winfo["seq_size"]
    for k in range(4):
        assert list(info.unexpected[k]) == winfo["unexpected"][k], ("unexpected", k)


def test_split_reference_suite_inputs(gpu, oracle):
    seen = set()
    for case in helpers.manifest("ref_suite"):
        key = (case["input"], tuple(case["ennaf_args"]))
        if key in seen:
            continue
        seen.add(key)
        kw = helpers.parse_ennaf_args(case["ennaf_args"])
        kw.pop("level", None)
        check_split(gpu, oracle, helpers.golden("ref_suite", case["input"]), **kw)


def test_split_cases(gpu, oracle):
    for case in helpers.manifest("cases"):
        kw = helpers.parse_ennaf_args(case["ennaf_args"])
        kw.pop("level", None); kw.pop("title", None)
        check_split(gpu, oracle, helpers.golden("cases", case["name"] + ".txt.gz"), **kw)


def _fuzz_fasta(rng):
    out = bytearray()
    if rng.random() < 0.3:
        out += rng.choice([b"\n", b" \n", b"\r\n\n", b"\t\n"])
    for _ in range(rng.randint(0, 6)):
        out += b">" + bytes(rng.choice(b"abcXYZ019|._\xc3\xfe") for _ in range(rng.randint(0, 8)))
        if rng.random() < 0.6:
            out += rng.choice([b" ", b"\t", b"  "]) + bytes(rng.choice(b"abc def\t\x02>@\xfe") for _ in range(rng.randint(0, 10)))
        out += rng.choice([b"\n", b"\r\n", b"\n\n", b"\x0b", b""])
        for _ in range(rng.randint(0, 4)):
            out += bytes(rng.choice(b"ACGTacgtNnRYKMSWBDHV-UuXxZ*> \t.12\xe0") for _ in range(rng.randint(0, 90)))
            out += rng.choice([b"\n", b"\r\n", b"\n\n", b"\n \n", b"\x0c", b""])
    return bytes(out)


def _fuzz_fastq(rng):
    out = bytearray()
    for _ in range(rng.randint(1, 6)):
        out += b"@" + bytes(rng.choice(b"abcXYZ019") for _ in range(rng.randint(0, 8)))
        if rng.random() < 0.6:
            out += b" " + bytes(rng.choice(b"abc def/12") for _ in range(rng.randint(0, 10)))
        out += b"\n"
        L = rng.randint(1, 100)
        s = bytes(rng.choice(b"ACGTacgtNnRY.x") for _ in range(L))
        if rng.random() < 0.2:
            s = s[:L // 2] + b" " + s[L // 2:]
        out += s + rng.choice([b"\n", b"\n\n"]) + b"+" + rng.choice([b"", b"xyz"]) + rng.choice([b"\n", b"\n\n"])
        q = bytes(rng.choice(b"!#$%IJK@+~") for _ in range(L))
        if rng.random() < 0.1:
            q = q[:L // 2] + b"\x01" + q[L // 2 + 1:]
        if rng.random() < 0.05:
            q = q[:-1]
        if rng.random() < 0.03:
            out += q + b"\nGARBAGE\n"
            continue
        out += q + rng.choice([b"\n", b"\n\n", b""])
    return bytes(out)


def test_split_fuzz_non_well_formed(gpu, oracle):
    """blank lines, CR/LF, tabs, unexpected bytes, '>' mid-line, truncated / malformed FASTQ (error strings too)"""
    rng = random.Random(1234)
    for it in range(300):
        text = _fuzz_fastq(rng) if rng.random() < 0.35 else _fuzz_fasta(rng)
        kw = {"seq_type": rng.choice(["dna", "rna", "protein", "text"])}
        if rng.random() < 0.25:
            kw["no_mask"] = True
        if rng.random() < 0.1:
            kw["strict"] = True
        check_split(gpu, oracle, text, **kw)


def test_split_tile_boundaries(gpu, oracle):
    """records and lines straddling the 64-byte thread chunks and the 16 KB tiles, headers longer than a tile"""
    rng = np.random.default_rng(5)
    parts = []
    for i in range(40):
        name = b"r%d " % i + bytes(rng.integers(97, 123, int(rng.integers(0, 40000)) if i % 7 == 0 else 5).astype(np.uint8))
        L = int(rng.integers(0, 70000))
        seq = bytes(np.frombuffer(b"ACGTacgtN", dtype=np.uint8)[rng.integers(0, 9, L)])
        w = int(rng.choice([0, 1, 60, 61, 63, 64, 65, 16384]))
        body = seq if w == 0 else b"\n".join(seq[k:k + w] for k in range(0, L, w))
        parts.append(b">" + name + b"\n" + body + (b"\n" if L else b""))
    check_split(gpu, oracle, b"".join(parts))
    check_split(gpu, oracle, b"".join(parts)[:-1])          # no final newline
    check_split(gpu, oracle, synth.fastq(5000, 151, seed=9), expect_fast=True)


def test_split_canonical_uses_fast_parser(gpu, oracle):
    """BASELINE-shaped inputs must be accepted by the canonical-input parser (and equal the oracle); one stray
    byte anywhere must send the whole input through the general parser with identical results"""
    rng = np.random.default_rng(17)
    fq = synth.fastq(30000, 150, seed=41)
    fa = synth.fasta_softmasked(3_000_000, width=60, seed=42, n_records=5, repeats=True, n_gaps=2)
    cases = [(fq, {}), (fq[:-1], {}), (fa, {}), (fa[:-1], {}), (synth.ont_fasta(40, 1000, 50000, seed=43), {}),
             (synth.protein_fasta(20000, 300, seed=44), {"seq_type": "protein"}), (synth.protein_fasta(2000, 300, seed=44), {"seq_type": "protein", "no_mask": True}),
             (synth.protein_fasta(2000, 300, seed=45), {"seq_type": "text"}), (synth.fastq(3000, 150, seed=46, lowercase=True, iupac=True), {}),
             (synth.fastq(3000, 150, seed=46, lowercase=True), {"no_mask": True}), (synth.fasta_reads(3000, 150, seed=47).replace(b"T", b"U"), {"seq_type": "rna"})]
    for L in (1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000):
        cases.append((b"".join(b"@q%d %d/1\n" % (i, i) + bytes(np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, L)]) + b"\n+\n" +
                               bytes(rng.integers(33, 127, L).astype(np.uint8)) + b"\n" for i in range(700)), {"_fast": L >= 31}))
    for w in (1, 2, 59, 63, 64, 65, 16383, 16384, 16385):
        s = bytes(np.frombuffer(b"ACGTacgtN-", dtype=np.uint8)[rng.integers(0, 10, 20 * w + 7)])
        cases.append((b"".join(b">r%d some comment here\n" % i + b"\n".join(s[k:k + w] for k in range(0, len(s), w)) + b"\n" for i in range(4)), {}))
    for text, kw in cases:
        kw = dict(kw)
        # lines of a few bytes break a 16 KB tile into more runs than the fast parser lists (condition C7): either parser is fine there
        check_split(gpu, oracle, text, expect_fast=True if kw.pop("_fast", True) else None, **kw)
    for text, at, byte in [(fq, 1_000_003, b"\r"), (fq, len(fq) - 5, b" "), (fa, 2_000_000, b"Z"), (fa, 17, b"\t"), (fa, 1_500_000, b"\x7f")]:
        check_split(gpu, oracle, text[:at] + byte + text[at + 1:], expect_fast=False)


def test_crlf_fasta_uses_fast_parser(gpu, oracle):
    """CR LF FASTA (process.c: '\\r' is an end-of-line byte, runs of them collapse) splits like LF FASTA in the single-pass
    transform -- no fallback -- and equals the oracle; line widths that put CR | LF across 16 KB tile boundaries; a stray CR
    still goes to the general parser"""
    rng = np.random.default_rng(23)
    fa = synth.fasta_softmasked(2_000_000, width=60, seed=42, n_records=5, repeats=True, n_gaps=2).replace(b"\n", b"\r\n")
    cases = [(fa, {}), (fa[:-1], {}), (fa[:-2], {}), (synth.ont_fasta(20, 1000, 50000, seed=43).replace(b"\n", b"\r\n"), {}),
             (synth.protein_fasta(5000, 300, seed=44).replace(b"\n", b"\r\n"), {"seq_type": "protein"}),
             (synth.protein_fasta(2000, 300, seed=45).replace(b"\n", b"\r\n"), {"seq_type": "text", "no_mask": True})]
    for w in (61, 62, 63, 126, 127, 254, 1022, 16382, 16383):
        s = bytes(np.frombuffer(b"ACGTacgtN-", dtype=np.uint8)[rng.integers(0, 10, 40 * w + 7)])
        cases.append((b"".join(b">r%d some comment here\r\n" % i + b"\r\n".join(s[k:k + w] for k in range(0, len(s), w)) + b"\r\n" for i in range(4)), {}))
    for text, kw in cases:
        check_split(gpu, oracle, text, expect_fast=not text.endswith(b"\r"), **kw)
    check_split(gpu, oracle, fa[:1_000_000] + b"\r" + fa[1_000_000:], expect_fast=False)
    naf = gpu.encode(fa)
    assert gpu.decode(naf) == fa.replace(b"\r\n", b"\n")


def test_zstd_compress_roundtrip(gpu, oracle):
    """our frames are valid zstd: the oracle decoder (pinned to libzstd) regenerates the input"""
    rng = np.random.default_rng(3)
    datasets = [b"", b"A", b"AB", b"A" * 70000, bytes(range(256)) * 300]
    for n in [2, 3, 15, 16, 17, 255, 256, 1023, 1024, 1025, 5000, 32767, 32768, 32769, 65535, 65536, 65537, 200000]:
        datasets.append(bytes(rng.choice(np.frombuffer(b"\x11\x12\x14\x18\x21\x22\x24\x28\x41\x42\x44\x48\x81\x82\x84\x88", dtype=np.uint8), n)))
        datasets.append((np.clip(np.round(rng.normal(34, 6, n)), 2, 40).astype(np.uint8) + 33).tobytes())
    datasets.append(b"".join(b"SRR1.%d\0" % i for i in range(30000)))
    datasets.append(np.tile(np.array([150, 0, 0, 0], dtype=np.uint8), 50000).tobytes())
    datasets.append(rng.integers(0, 256, 300000, dtype=np.uint8).tobytes())
    datasets.append(bytes(rng.choice(np.arange(200, dtype=np.uint8), 100000, p=np.r_[[0.5], np.full(199, 0.5 / 199)])))
    datasets.append(b"".join(b"%d/1\0" % i for i in range(1, 40000)))
    datasets.append(b"".join(bytes([65 + (i * i) % 23]) * (1 + i % 40) for i in range(6000)))      # runs: offset-1 matches, literal length 0
    datasets.append(b"abcdefgh" * 9000 + b"tail")
    for level in (1, 3):                      # 1: Huffman-only 32 KB blocks; >= 2: 8 KB blocks with LZ77 matches + FSE-coded sequences
        for d in datasets:
            z = gpu.zstd_compress(d, level=level)
            assert z[:4] == b"\x28\xb5\x2f\xfd"
            assert oracle.zstd_decompress(z) == d, (level, len(d))
            assert gpu.zstd_decompress(z) == d, (level, len(d))
    ids = datasets[-7]
    assert ids.startswith(b"SRR1.0")
    assert len(gpu.zstd_compress(ids, level=3)) < 0.4 * len(gpu.zstd_compress(ids, level=1))


def test_lz_frames_equal_cpu_emulation(gpu, tmp_path):
    """the GPU runs the same HD bodies the not-gpu tests run on the CPU -- the data-parallel stage a level >= 2 selects
    (tests/emu/emu_zlzc.cpp) and the thread-per-block one behind NAFGPU_LZ=1 (tests/emu/emu_zenc.cpp): same bytes out"""
    import shutil
    import subprocess
    exes = {}
    for name in ("emu_zlzc", "emu_zenc"):
        exe = os.path.join(helpers.ROOT, "tests", "_build", name + ("_bytes" if name == "emu_zlzc" else ""))
        if shutil.which("g++"):                              # build from the sources of this snapshot; else the binary that travelled
            exe = str(tmp_path / name)
            subprocess.run(["g++", "-std=c++17", "-O2", "-DZLC_BYTES", "-o", exe, os.path.join(helpers.ROOT, "tests", "emu", name + ".cpp")], check=True)
        elif not os.path.exists(exe):
            pytest.skip("no %s binary and no g++ on this box" % name)
        exes[name] = exe
    rng = np.random.default_rng(9)
    saved = os.environ.get("NAFGPU_LZ")
    try:
        for d in [b"".join(b"SRR1.%d\0" % i for i in range(70000, 90000)), b"".join(b"%d/1\0" % i for i in range(1, 30000)),
                  np.tile(np.array([150, 0, 0, 0], dtype=np.uint8), 30000).tobytes(), bytes(rng.integers(0, 5, 100000, dtype=np.uint8)),
                  b"abcdefgh" * 9000 + b"tail", b"", b"q" * 20000, b"r%d\0" % 7 * 5000 + bytes(rng.integers(0, 256, 9000, dtype=np.uint8))]:
            inp, z = str(tmp_path / "i.bin"), str(tmp_path / "o.zst")
            with open(inp, "wb") as f:
                f.write(d)
            os.environ.pop("NAFGPU_LZ", None)
            assert subprocess.run([exes["emu_zlzc"], inp, z, "8192"], capture_output=True).returncode == 0
            assert gpu.zstd_compress(d, level=3) == open(z, "rb").read(), ("data-parallel stage", len(d))
            os.environ["NAFGPU_LZ"] = "1"                    # (the library reads the switch per call)
            assert subprocess.run([exes["emu_zenc"], inp, z, "8192", "1", "32"], capture_output=True).returncode == 0
            assert gpu.zstd_compress(d, level=3) == open(z, "rb").read(), ("thread-per-block stage", len(d))
    finally:
        if saved is None:
            os.environ.pop("NAFGPU_LZ", None)
        else:
            os.environ["NAFGPU_LZ"] = saved


def test_levels_lz_on_text_streams(gpu, oracle):
    """ennaf -# : level >= 2 parses ids / comments / lengths with matches, level 1 (the default) does not; both files
    are valid for every decoder, and the sequence / quality streams do not depend on the level"""
    for text, kw in [(synth.fastq(60_000, 150, seed=11), {}), (synth.ont_fasta(200, 10000, 30000, seed=12), {}),
                     (synth.protein_fasta(30_000, 300, seed=13), {"seq_type": "protein"})]:
        naf0, i0 = gpu.encode_with_info(text, level=1, **kw)
        naf1, i1 = gpu.encode_with_info(text, level=3, **kw)
        for naf in (naf0, naf1):
            assert gpu.decode(naf) == text
            assert oracle.decode(naf) == text
            if helpers.have_ref():
                rc, out, err = helpers.ref_run("unnaf", [], naf)
                assert rc == 0 and out == text, err
        assert list(i0.stream_raw) == list(i1.stream_raw)
        assert i0.stream_comp[4] == i1.stream_comp[4] and i0.stream_comp[5] == i1.stream_comp[5]
        assert i1.stream_comp[0] + i1.stream_comp[1] < 0.75 * (i0.stream_comp[0] + i0.stream_comp[1])
        # (the whole file need not shrink: a mask stream without repeats pays for its smaller blocks -- measured +114 B on the ONT case)
        assert len(naf1) < len(naf0) + 4096


def test_encode_reference_suite(gpu, oracle):
    """the reference's 60 suite cases with OUR ennaf in the pipeline: oracle unnaf (and the reference unnaf when
    built) must print the pinned stdout; our stderr report must equal the pinned ennaf stderr."""
    from naf_b200 import cli_text
    for case in helpers.manifest("ref_suite"):
        text = helpers.golden("ref_suite", case["input"])
        ekw, ukw = helpers.parse_ennaf_args(case["ennaf_args"]), helpers.parse_unnaf_args(case["unnaf_args"])
        expect = helpers.golden("ref_suite", case["set"], case["name"] + ".out")
        naf, info = gpu.encode_with_info(text, **ekw)
        assert cli_text.unexpected_report(info, ekw.get("seq_type", "dna")) == helpers.golden("ref_suite", case["set"], case["name"] + ".e.err"), case["name"]
        assert oracle.decode(naf, **ukw) == expect, case["name"]
        assert gpu.unnaf(naf, **ukw) == expect, case["name"]
        if helpers.have_ref():
            rc, out, err = helpers.ref_run("unnaf", case["unnaf_args"], naf)
            assert rc == 0 and out == expect, (case["name"], err)


def test_encode_cases_decoded_by_reference(gpu, oracle):
    for case in helpers.manifest("cases"):
        text = helpers.golden("cases", case["name"] + ".txt.gz")
        ekw = helpers.parse_ennaf_args(case["ennaf_args"])
        naf = gpu.encode(text, **ekw)
        for key in ("fasta", "fastq", "seq", "sequences", "4bit", "ids", "names", "lengths", "mask", "charcount", "title", "number"):
            exp = case["views"][key]
            if exp["rc"] != 0:
                continue
            got = oracle.decode(naf, key)
            assert (len(got), helpers.sha(got)) == (exp["size"], exp["sha256"]), (case["name"], key)
            if helpers.have_ref():
                rc, out, err = helpers.ref_run("unnaf", ["--" + key], naf)
                assert rc == 0 and helpers.sha(out) == exp["sha256"], (case["name"], key, err)


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref binaries not built")
def test_encode_medium_configs_decoded_by_reference(gpu):
    """config-2 / 3 / 4 / 5 shapes at sizes the reference decodes in seconds"""
    for text, kw, args in [
        (synth.fastq(200_000, 150, seed=31), {}, []),
        (synth.ont_fasta(300, 10000, 50000, seed=32), {}, []),
        (synth.protein_fasta(100_000, 300, seed=33), {"seq_type": "protein"}, []),
        (synth.fasta_softmasked(40_000_000, 60, seed=34, n_records=2, repeats=True, n_gaps=3), {}, []),
    ]:
        naf = gpu.encode(text, **kw)
        rc, out, err = helpers.ref_run("unnaf", args, naf)
        assert rc == 0, err
        assert out == text
        assert gpu.decode(naf) == text


def test_error_messages_match_reference_strings(gpu, oracle):
    for text, kw in [(b"ACGT\n", {}), (b"x>a\nAC\n", {}), (b"@r\nACGT\n+\nII\n", {}), (b"@r\nACGT\n", {}), (b"@r", {}),
                     (b"@r\nAC\nXX\nII\n", {}), (b"@r\nAC\n+\nII\nzz\n", {}), (b">a\nACZT\n", {"strict": True})]:
        check_split(gpu, oracle, text, **kw)


def test_roundtrip_tiny_records(gpu, oracle):
    """thousands of records per 16 KB of text: more records than a tile of the text writer stages in shared memory, more
    runs than the fast parser lists per tile -- both directions must still be exact"""
    rng = np.random.default_rng(71)
    texts = []
    for L in (1, 2, 5, 31):
        texts.append(b"".join(b"@%d\n" % i + bytes(np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, L)]) + b"\n+\n" +
                              bytes(rng.integers(33, 127, L).astype(np.uint8)) + b"\n" for i in range(20000)))
    texts.append(b"".join(b">s%d\n" % i + b"\n".join(bytes(np.frombuffer(b"ACGTacgt", dtype=np.uint8)[rng.integers(0, 8, 3)]) for _ in range(4)) + b"\n" for i in range(20000)))
    texts.append(b"".join(b">%d\n" % i for i in range(50000)))                      # names only, no sequence at all
    for text in texts:
        naf = gpu.encode(text)
        want = oracle.decode(oracle.encode(text)[0])
        assert gpu.decode(naf) == want
        assert oracle.decode(naf) == want
        assert gpu.decode(oracle.encode(text)[0]) == want
        for view in ("fasta", "ids", "sequences", "seq"):
            assert gpu.decode(naf, view) == oracle.decode(naf, view), view
