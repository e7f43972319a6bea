"""GPU parity at BASELINE.json's full sizes, through size-independent properties (the oracle needs minutes at these sizes):
decode(encode(x)) == x through the host-buffer C ABI, stream sizes consistent with the container header, and the
canonical-input parser accepted the input.  Inputs are the synthetic read sets of SURVEY 8(d) (fixed seeds)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers
from naf_b200 import container, synth

pytestmark = pytest.mark.gpu


def _workdir(tmp_path, need_bytes):
    """tmpfs if it has room for the files of a cross-check (text + .naf + text again), else pytest's tmp dir"""
    try:
        if shutil.disk_usage("/dev/shm").free > need_bytes + (2 << 30):
            d = os.path.join("/dev/shm", f"naf_full_{os.getpid()}")
            os.makedirs(d, exist_ok=True)
            return d
    except OSError:
        pass
    return str(tmp_path)


def _cross_check_with_reference(gpu, text, naf, tmp_path, ennaf_args=()):
    """both directions of the drop-in claim at full size: the unmodified unnaf decodes OUR file to the input, and WE decode
    the unmodified ennaf's file to the input (files, not pipes: the tools stream them)"""
    if not helpers.have_ref():
        return
    d = _workdir(tmp_path, 2 * len(text) + 2 * len(naf))
    try:
        fin, fnaf, fout, fref = (os.path.join(d, x) for x in ("in.txt", "ours.naf", "out.txt", "ref.naf"))
        env = dict(os.environ, TMPDIR=d)
        with open(fnaf, "wb") as f:
            f.write(naf)
        subprocess.run([os.path.join(helpers.REF_BIN, "unnaf"), fnaf, "-o", fout], check=True, env=env)
        with open(fout, "rb") as f:
            assert f.read() == text, "reference unnaf on our .naf differs from the input"
        os.remove(fout); os.remove(fnaf)
        with open(fin, "wb") as f:
            f.write(text)
        subprocess.run([os.path.join(helpers.REF_BIN, "ennaf"), *ennaf_args, fin, "-o", fref], check=True, env=env)
        os.remove(fin)
        with open(fref, "rb") as f:
            ref_naf = f.read()
        assert gpu.decode(ref_naf) == text, "our decode of the reference-made .naf differs from the input"
    finally:
        if d != str(tmp_path):
            shutil.rmtree(d, ignore_errors=True)


def _roundtrip(gpu, text, n_records, n_bases, tmp_path=None, **kw):
    naf, info = gpu.encode_with_info(text, **kw)
    assert gpu.timing().parser_fallback == 0
    assert info.n_sequences == n_records and info.n_bases == n_bases
    h = container.read_header(naf)
    assert h.n_sequences == n_records
    out = gpu.decode(naf)
    assert len(out) == len(text) and out == text
    del out
    if tmp_path is not None:
        _cross_check_with_reference(gpu, text, naf, tmp_path, ["--" + kw["seq_type"]] if "seq_type" in kw else [])
    return len(naf) / len(text)


def test_config2_fastq_10M_reads(gpu, tmp_path):
    text = synth.fastq(10_000_000, 150, seed=42)
    ratio = _roundtrip(gpu, text, 10_000_000, 1_500_000_000, tmp_path)
    assert 0.30 < ratio < 0.45


def test_config4_protein_1M(gpu, tmp_path):
    text = synth.protein_fasta(1_000_000, 300, seed=42)
    ratio = _roundtrip(gpu, text, 1_000_000, 300_000_000, tmp_path, seq_type="protein")
    assert 0.45 < ratio < 0.62


def test_config3_ont_like_fasta(gpu):
    text = synth.ont_fasta(100_000, 10000, 50000, seed=42)
    naf, info = gpu.encode_with_info(text)
    assert gpu.timing().parser_fallback == 0 and info.n_sequences == 100_000
    out = gpu.decode(naf)
    assert out == text
    # the other views agree with each other at this size: --seq is the concatenation of --sequences lines
    seq = gpu.decode(naf, "seq")
    assert len(seq) == info.n_bases
    assert gpu.decode(naf, "sequences").replace(b"\n", b"") == seq


def test_config5_3gbp_softmasked_fasta(gpu, tmp_path):
    text = synth.fasta_softmasked(3_000_000_000, 60, seed=42, n_records=24, repeats=True, n_gaps=20)
    ratio = _roundtrip(gpu, text, 24, 3_000_000_000, tmp_path)
    assert 0.2 < ratio < 0.3


def test_one_record_beyond_4_gib(gpu, tmp_path):
    """more than 2^32 bases in ONE record, more than 4 GiB of text in one call: every 32-bit offset on the path would wrap,
    and the length needs a continuation unit (ennaf/src/encoders.c:78-87, unnaf/src/output.c:390-393).  Checked against the
    unmodified unnaf as well (it streams the file in ~10 s)."""
    n = (1 << 32) + 123_457
    rng = np.random.default_rng(5)
    block = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 1 << 26)]
    block[5_000_000:5_300_000] |= 0x20                                   # soft-masked runs, a few IUPAC codes
    block[40_000_000:40_000_700] |= 0x20
    block[100::1_000_003] = ord("N")
    block[7::9_000_011] = ord("r")
    head = b">big one record\n"
    text = np.empty(len(head) + n + 1, dtype=np.uint8)
    text[:len(head)] = np.frombuffer(head, dtype=np.uint8)
    body = text[len(head):len(head) + n]
    for at in range(0, n, block.size):
        m = min(block.size, n - at)
        body[at:at + m] = block[:m]
    body[(1 << 32) - 3:(1 << 32) + 5] = np.frombuffer(b"acgtACGT", dtype=np.uint8)      # something recognisable across the 2^32 line
    text[-1] = 10
    naf, info = gpu.encode_with_info(text)
    assert info.n_sequences == 1 and info.n_bases == n and info.longest_line == n
    lengths = np.frombuffer(gpu.decode(naf, "lengths"), dtype="<u4")
    assert lengths.tolist() == [0xFFFFFFFF, n - 0xFFFFFFFF]
    out = np.frombuffer(gpu.decode(naf), dtype=np.uint8)
    assert out.size == text.size and np.array_equal(out, text)
    del out
    if helpers.have_ref():
        d = _workdir(tmp_path, text.size + len(naf))
        try:
            fnaf, fout = os.path.join(d, "big.naf"), os.path.join(d, "big.fa")
            with open(fnaf, "wb") as f:
                f.write(naf)
            subprocess.run([os.path.join(helpers.REF_BIN, "unnaf"), fnaf, "-o", fout], check=True, env=dict(os.environ, TMPDIR=d))
            got = np.fromfile(fout, dtype=np.uint8)
            assert got.size == text.size and np.array_equal(got, text), "reference unnaf on our > 4 GiB file differs from the input"
        finally:
            if d != str(tmp_path):
                shutil.rmtree(d, ignore_errors=True)


def test_config2_level2_decoded_by_reference(gpu, tmp_path):
    """ennaf -2 at the full size: names and lengths go through the data-parallel LZ stage (csrc/zstd_lzc_hd.cuh: 32 k blocks with
    per-stream tables).  The file is smaller than level 1's, we decode it back, and so does the unmodified reference unnaf."""
    text = synth.fastq(10_000_000, 150, seed=42)
    naf1 = gpu.encode(text)
    n1 = len(naf1)
    del naf1
    naf2, info = gpu.encode_with_info(text, level=2)
    assert info.n_sequences == 10_000_000 and gpu.timing().parser_fallback == 0
    assert len(naf2) < n1 and 0.34 < len(naf2) / len(text) < 0.37
    assert gpu.decode(naf2) == text
    if helpers.have_ref():
        d = _workdir(tmp_path, len(text) + len(naf2))
        try:
            fnaf, fout = os.path.join(d, "l2.naf"), os.path.join(d, "l2.txt")
            with open(fnaf, "wb") as f:
                f.write(naf2)
            subprocess.run([os.path.join(helpers.REF_BIN, "unnaf"), fnaf, "-o", fout], check=True, env=dict(os.environ, TMPDIR=d))
            with open(fout, "rb") as f:
                assert f.read() == text, "reference unnaf on our level-2 .naf differs from the input"
        finally:
            if d != str(tmp_path):
                shutil.rmtree(d, ignore_errors=True)
