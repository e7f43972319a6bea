"""GPU: the host-buffer calls overlap their copies with the kernels (HostPipe, csrc/common.cuh): the text goes up in
chunks with one transform launch per chunk, the .naf goes up in chunks with the last big stream decoded piece by piece,
and finished pieces of the text go down while the next ones are produced.  The counterpart in the reference is its
streaming through 16 KB / 128 KB windows (ennaf/src/process.c:227-240, unnaf/src/output.c:640-650).  Results must be
byte-identical to the unpiped calls, to the oracle and to the reference; thresholds are lowered through the
environment so that files of a few MB take the piped paths."""
import os

import pytest

import helpers
from naf_b200 import synth

pytestmark = pytest.mark.gpu

SMALL = {"NAFGPU_PIPE_MIN": str(1 << 20), "NAFGPU_PIPE_CHUNK": str(1 << 20), "NAFGPU_PIPE_PIECE": str(1 << 18)}


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def texts():
    yield "fastq", synth.fastq(60_000, 150, seed=31), {}
    yield "fastq-iupac", synth.fastq(40_000, 151, seed=32, iupac=True), {}
    yield "fasta-softmasked", synth.fasta_softmasked(20_000_000, 60, seed=33, n_records=5, repeats=True, n_gaps=2), {}
    yield "ont", synth.ont_fasta(300, 10000, 50000, seed=34), {}
    yield "protein", synth.protein_fasta(40_000, 300, seed=35), {"seq_type": "protein"}


def test_piped_calls_equal_unpiped_calls(gpu, oracle):
    for name, text, kw in texts():
        with env(NAFGPU_PIPE="0"):
            naf0 = gpu.encode(text, **kw)
            out0 = gpu.decode(naf0)
        assert out0 == text, name
        with env(**SMALL):
            naf1 = gpu.encode(text, **kw)
            assert naf1 == naf0, name                       # same kernels, same order of tiles: the same file
            assert gpu.decode(naf0) == text, name
            assert gpu.timing().kernel_launches > 0
            for view in ("fasta", "sequences", "seq", "ids", "4bit" if "seq_type" not in kw else "names"):
                with env(NAFGPU_PIPE="0"):
                    want = gpu.decode(naf0, view)
                assert gpu.decode(naf0, view) == want, (name, view)
            n = 1 + text.count(b"\n>") if text[:1] == b">" else text.count(b"\n") // 4
            a = gpu.decode(naf0, first_record=0, n_records=n // 3)
            b = gpu.decode(naf0, first_record=n // 3, n_records=n - n // 3)
            assert a + b == text, name
        assert oracle.decode(naf1) == text, name


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref binaries not built")
def test_piped_decode_of_reference_made_files(gpu, tmp_path):
    """frames with sequences, inherited tables and matches across blocks cannot be cut: decoded whole, text still sent down in pieces"""
    for name, text, kw in texts():
        args = ["--" + kw["seq_type"]] if kw else []
        rc, naf, err = helpers.ref_run("ennaf", args + ["-c"], text, tmp=str(tmp_path))
        assert rc == 0, err
        with env(**SMALL):
            assert gpu.decode(naf) == text, name


def test_piped_encode_falls_back_to_the_general_parser(gpu, oracle):
    """a stray CR in the middle of a sequence line is not canonical (to process.c it is a line end): the piped call notices, waits
    for the whole upload and redoes the split.  CR LF line ends, on the other hand, stay on the single-pass transform."""
    plain = synth.fasta_softmasked(4_000_000, 60, seed=36, n_records=7)
    at = 2_000_000
    while not (plain[at - 1:at + 1].isalpha()):
        at += 1
    text = plain[:at] + b"\r" + plain[at:]
    with env(**SMALL):
        naf = gpu.encode(text)
        assert gpu.timing().parser_fallback == 1
        assert oracle.decode(naf) == plain
        naf = gpu.encode(plain.replace(b"\n", b"\r\n"))
        assert gpu.timing().parser_fallback == 0
    assert oracle.decode(naf) == plain


def test_piped_errors_leave_the_context_usable(gpu):
    import naf_b200
    text = synth.fastq(30_000, 150, seed=37)
    bad = text[:len(text) // 2] + b"@broken\nACGT\n+\nII\n" + text[len(text) // 2:]
    with env(**SMALL):
        naf = gpu.encode(text)
        with pytest.raises(naf_b200.NafGpuError):
            gpu.encode(bad)
        assert gpu.encode(text) == naf
        cut = bytearray(naf)
        cut[len(cut) - 300000] ^= 0xFF
        try:
            gpu.decode(bytes(cut))
        except naf_b200.NafGpuError:
            pass
        with pytest.raises(naf_b200.NafGpuError):
            gpu.decode(naf[:len(naf) - 1000])
        assert gpu.decode(naf) == text


def test_streamed_calls_equal_the_one_shot_calls(gpu, oracle):
    """nafgpu_encode_begin / _buffer / _feed / _end and nafgpu_decode_to (what bin/ennaf and bin/unnaf call): text handed over and
    received in pieces, same bytes as the one-shot calls; also when the total size was not announced (a pipe)"""
    for name, text, kw in texts():
        naf = gpu.encode(text, **kw)
        for hint in (len(text), 0):
            pieces = [text[i:i + 5_000_011] for i in range(0, len(text), 5_000_011)]
            assert gpu.encode_pieces(pieces, size_hint=hint, **kw) == naf, (name, hint)
            got = []
            gpu.encode_pieces(pieces, size_hint=hint, write=got.append, **kw)
            assert b"".join(got) == naf, (name, hint)
        for view in ("default", "fasta", "seq", "ids"):
            got = []
            total = gpu.decode_to(naf, got.append, view)
            want = gpu.decode(naf, view)
            assert b"".join(got) == want and total == len(want), (name, view)
        with env(**SMALL):
            got = []
            gpu.decode_to(naf, got.append)
            assert b"".join(got) == text, name
    assert gpu.encode_pieces([]) == gpu.encode(b"")
