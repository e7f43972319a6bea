"""GPU parity at BASELINE.json's full sizes, through size-independent properties (the oracle needs minutes at these sizes):
decode(encode(x)) == x through the host-buffer C ABI, stream sizes consistent with the container header, and the
canonical-input parser accepted the input.  Inputs are the synthetic read sets of SURVEY 8(d) (fixed seeds)."""
import numpy as np
import pytest

from naf_b200 import container, synth

pytestmark = pytest.mark.gpu


def _roundtrip(gpu, text, n_records, n_bases, **kw):
    naf, info = gpu.encode_with_info(text, **kw)
    assert gpu.timing().parser_fallback == 0
    assert info.n_sequences == n_records and info.n_bases == n_bases
    h = container.read_header(naf)
    assert h.n_sequences == n_records
    out = gpu.decode(naf)
    assert len(out) == len(text) and out == text
    return len(naf) / len(text)


def test_config2_fastq_10M_reads(gpu):
    text = synth.fastq(10_000_000, 150, seed=42)
    ratio = _roundtrip(gpu, text, 10_000_000, 1_500_000_000)
    assert 0.30 < ratio < 0.45


def test_config4_protein_1M(gpu):
    text = synth.protein_fasta(1_000_000, 300, seed=42)
    ratio = _roundtrip(gpu, text, 1_000_000, 300_000_000, seq_type="protein")
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


def test_config5_3gbp_softmasked_fasta(gpu):
    text = synth.fasta_softmasked(3_000_000_000, 60, seed=42, n_records=24, repeats=True, n_gaps=20)
    ratio = _roundtrip(gpu, text, 24, 3_000_000_000)
    assert 0.2 < ratio < 0.3
