"""GPU parity: decode of REFERENCE-produced .naf files must be byte-identical to the reference unnaf
(north_star: "our unnaf on a reference-produced .naf yields byte-identical FASTA/FASTQ").
Everything goes through the C ABI (naf_b200 -> libnafgpu.so); expected bytes come from the committed
goldens (made by the unmodified reference, tools/make_golden.py) and from the oracle."""
import os

import numpy as np
import pytest

import helpers
from naf_b200 import synth

pytestmark = pytest.mark.gpu


def test_zstd_frames_golden(gpu, oracle):
    """every committed frame (libzstd levels -5..19, decodecorpus: all block/literal/sequence modes)"""
    n = 0
    for e in helpers.manifest("zstd"):
        z = helpers.golden("zstd", e["frame"])
        out = gpu.zstd_decompress(z)
        assert len(out) == e["size"], e["frame"]
        assert helpers.sha(out) == e["sha256"], e["frame"]
        assert out == oracle.zstd_decompress(z), e["frame"]
        n += 1
    assert n >= 100


def test_ref_suite_decode(gpu):
    """the reference's own perl-suite cases: decode the reference-made .naf with the same unnaf flags"""
    for case in helpers.manifest("ref_suite"):
        naf = helpers.golden("ref_suite", case["naf"])
        kw = helpers.parse_unnaf_args(case["unnaf_args"])
        expect = helpers.golden("ref_suite", case["set"], case["name"] + ".out")
        got = gpu.unnaf(naf, **kw)
        assert got == expect, (case["name"], got[:200], expect[:200])


def test_cases_all_views(gpu, oracle):
    """FASTQ / RNA / protein / text / mask edge cases / multi-block streams at levels 1,3,19 / --long 31"""
    for case in helpers.manifest("cases"):
        naf = helpers.golden("cases", case["name"] + ".naf")
        for key, exp in case["views"].items():
            parts = key.split()
            kw = helpers.parse_unnaf_args(["--" + parts[0]] + parts[1:])
            if exp["rc"] != 0:
                with pytest.raises(Exception):
                    gpu.unnaf(naf, **kw)
                continue
            got = gpu.unnaf(naf, **kw)
            assert len(got) == exp["size"], (case["name"], key)
            assert helpers.sha(got) == exp["sha256"], (case["name"], key)


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref binaries not built")
def test_medium_fastq_vs_reference(gpu, tmp_path):
    """200k x 150 bp FASTQ (config-2 shape, multi-block every stream): reference ennaf -> our decode"""
    text = synth.fastq(200_000, 150, seed=21)
    rc, naf, err = helpers.ref_run("ennaf", ["-c"], text, tmp=str(tmp_path))
    assert rc == 0, err
    assert gpu.decode(naf) == text


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("level", ["-1", "-9"])
def test_medium_fasta_vs_reference(gpu, tmp_path, level):
    """30 Mbp soft-masked FASTA with repeats: LZ sequences, treeless literals, Repeat_Mode"""
    text = synth.fasta_softmasked(30_000_000, width=60, seed=22, n_records=3, repeats=True, n_gaps=2)
    rc, naf, err = helpers.ref_run("ennaf", [level, "-c"], text, tmp=str(tmp_path))
    assert rc == 0, err
    assert gpu.decode(naf) == text
    rc, seq, _ = helpers.ref_run("unnaf", ["--seq", "--no-mask"], naf)
    assert gpu.decode(naf, "seq", no_mask=True) == seq


def test_oracle_roundtrip_property(gpu, oracle):
    """size-independent property: oracle encode (raw-block frames) -> GPU decode == input, ONT-like config-3 shape"""
    text = synth.ont_fasta(40, 10000, 50000, seed=23)
    naf, _ = oracle.encode(text)
    assert gpu.decode(naf) == text
    assert gpu.decode(naf, "fasta", line_length=0) == oracle.decode(naf, "fasta", line_length=0)


def test_corrupt_input_never_crashes(gpu, oracle, tmp_path):
    """bit flips, truncation and overwritten spans in .naf files made by us, by the oracle and by the reference: every
    call returns (an error like the reference's die(), or some text -- zstd frames here carry no checksum), nothing hangs,
    and the context keeps working (cf. unnaf/src/utils.c:52 "incomplete or truncated input")"""
    import random
    import naf_b200
    rng = random.Random(7)
    texts = [synth.fastq(3000, 150, seed=1, lowercase=True), synth.fasta_softmasked(300000, 60, seed=2, n_records=3, repeats=True)]
    files = []
    for t in texts:
        files += [gpu.encode(t), gpu.encode(t, level=3), oracle.encode(t)[0]]
        if helpers.have_ref():
            rc, naf, err = helpers.ref_run("ennaf", ["-c"], t, tmp=str(tmp_path))
            assert rc == 0, err
            files.append(naf)
    errors = 0
    for it in range(300):
        naf = bytearray(rng.choice(files))
        kind = rng.random()
        if kind < 0.5:
            for _ in range(rng.randint(1, 3)):
                naf[rng.randrange(len(naf))] ^= 1 << rng.randrange(8)
        elif kind < 0.75:
            naf = naf[:rng.randrange(1, len(naf))]
        else:
            at = rng.randrange(len(naf))
            naf[at:at + rng.randint(1, 64)] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 64)))
        try:
            gpu.decode(bytes(naf))
        except naf_b200.NafGpuError:
            errors += 1
    assert errors > 50
    for f in files[:3]:
        assert gpu.decode(f) == oracle.decode(f)


def test_absurd_sizes_fail_cleanly_and_leave_the_context_usable(gpu, oracle):
    """a header that claims terabytes (one flipped VLE byte) is refused like any damaged file -- and the calls after it work
    (a failed giant allocation used to stay in the arena's high-water mark and made every second call fail)"""
    import naf_b200
    text = synth.fastq(2000, 150, seed=5)
    naf = gpu.encode(text)
    bad = helpers.claim_huge_ids(naf)
    for _ in range(3):
        with pytest.raises(naf_b200.NafGpuError):
            gpu.decode(bad)
        assert gpu.decode(naf) == text
        assert gpu.encode(text) == naf


def test_stated_limits_are_refused_not_wrapped(gpu):
    """An ids stream of 4 GiB (terminator offsets are 32-bit) is refused with NAFGPU_E_UNSUPPORTED and a message that says so --
    not decoded with wrapped offsets.  The file is 130 KB: a frame of 32,768 RLE blocks of 128 KB of zeros."""
    import naf_b200
    from naf_b200.container import put_vle
    nblk = 32768
    frame = bytes([0x00, (17 - 10) << 3])
    for i in range(nblk):
        bh = (1 if i == nblk - 1 else 0) | (1 << 1) | (131072 << 3)
        frame += bytes([bh & 0xFF, (bh >> 8) & 0xFF, (bh >> 16) & 0xFF, 0])
    naf = bytes([0x01, 0xF9, 0xEC, 1, 1 << 5, ord(" ")]) + put_vle(0) + put_vle(1 << 32) + put_vle(1 << 32) + put_vle(len(frame)) + frame
    with pytest.raises(naf_b200.NafGpuError) as e:
        gpu.decode(naf, "ids")
    assert e.value.code == -4 and "4 GiB" in e.value.message, (e.value.code, e.value.message)
    text = synth.fastq(500, 150, seed=6)
    assert gpu.decode(gpu.encode(text)) == text


def test_fasta_surplus_sequence_is_printed_like_the_reference(gpu, oracle, tmp_path):
    """ennaf's id-byte bug (an unexpected byte in an id puts its replacement into the SEQUENCE) makes files whose
    sequence is longer than their lengths add up to; unnaf prints the surplus after the last record, into what is left of its
    last line, without a final newline (output.c:420-427).  Same bytes from us, for every maker of the file."""
    cases = [(b">>\na", "text"), (b">a\x01b\nACGTAC\nGG\n>c\x02\x03\nTTTTTTT\n", "dna"),
             (b">a\x01b x\nacgtnnac\nGG\n>c\x02\x03\nTTTTTTT\n>e\n>f\n", "dna"), (b">p\x01\nMKV\nLLA\n>q\x7f\x01\x01\nMM\n", "protein"),
             (b"@r\x01\nACGT\n+\nIIII\n", "dna"), (b">u\x01\x01\x01\nACGU\n" + b">v\x02\nacguACGU\n" * 300, "rna")]
    for text, st in cases:
        files = [oracle.encode(text, seq_type=st)[0], gpu.encode(text, seq_type=st), gpu.encode(text, seq_type=st, level=3)]
        if helpers.have_ref():
            rc, naf, err = helpers.ref_run("ennaf", ["--" + st, "-c"], text, tmp=str(tmp_path))
            assert rc == 0, err
            files.append(naf)
        for naf in files:
            for kw, args in (({}, []), ({"line_length": 3}, ["--line-length", "3"]), ({"line_length": 0}, ["--line-length", "0"]),
                             ({"no_mask": True}, ["--no-mask"])):
                want = oracle.decode(naf, "fasta", **kw)
                assert gpu.decode(naf, "fasta", **kw) == want, (text, st, kw)
                if helpers.have_ref():
                    rc, out, err = helpers.ref_run("unnaf", ["--fasta"] + args, naf, timeout=20)
                    assert rc == 0 and out == want, (text, st, kw, err)


def test_damaged_headers_that_used_to_read_out_of_bounds(gpu, oracle):
    """(a) a FASTQ file whose quality stream is shorter than its lengths add up to, (b) a record count far beyond the
    number of '\\0' terminators in the ids (2^62: the byte size of the offset array wrapped), (c) a ranged decode of (a):
    all refused as damaged files, and the context keeps working"""
    import naf_b200
    long_reads, short_reads = synth.fastq(3000, 150, seed=3), synth.fastq(3000, 100, seed=3)
    naf, donor = gpu.encode(long_reads), gpu.encode(short_reads)
    bad_q = helpers.replace_section(naf, 5, donor)
    for kw in ({}, {"first_record": 2000, "n_records": 1000}):
        with pytest.raises(naf_b200.NafGpuError) as e:
            gpu.decode(bad_q, "fastq", **kw)
        assert e.value.code == -3
    assert gpu.decode(bad_q, "fasta") == oracle.decode(naf, "fasta")      # FASTA output does not touch the quality
    for n in (3001, 1 << 40, 1 << 62, (1 << 64) - 1):
        with pytest.raises(naf_b200.NafGpuError) as e:
            gpu.decode(helpers.claim_records(naf, n))
        assert e.value.code == -3, n
        assert gpu.decode(naf) == long_reads


def test_damaged_block_index_falls_back_to_the_header_walk(gpu, oracle):
    """our files carry a block index (a zstd skippable frame behind the lengths frame, which the reference skips); a file
    whose index is damaged but whose streams are fine must still decode -- by the walk -- and a file without one as well"""
    from naf_b200 import container
    text = synth.fastq(40_000, 150, seed=41)
    naf = gpu.encode(text)
    plain = gpu.encode(text, block_index=False)
    assert len(naf) > len(plain) and gpu.decode(plain) == text
    at = naf.find(b"NAFGIDX1")
    assert at > 0 and plain.find(b"NAFGIDX1") < 0
    h = container.read_header(naf)
    orig, comp, off = h.sections[2]
    assert off < at < off + comp                                    # inside the lengths section
    for delta in (40, 41, 60, 200):                                 # bytes of the entry table / of the size arrays
        bad = bytearray(naf)
        bad[at + delta] ^= 0x15
        assert gpu.decode(bytes(bad)) == text, delta
        assert oracle.decode(bytes(bad)) == text
    assert oracle.decode(naf) == text
