"""not-gpu: the decoder's HD kernel bodies (naf_b200/csrc/zstd_dec.cuh), run through the test-only host
executor in tests/emu, against the golden frames.  This checks the device logic of the thread-serial
kernels (headers, Huffman/FSE tables, sequence decode, repeat-offset scan, pointer jumping) on the CPU;
the shipped library never contains this executor."""
import os
import subprocess

import pytest

import helpers

ROOT = helpers.ROOT
EXE = os.path.join(ROOT, "tests", "_build", "emu_zstd")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "emu", "emu_zstd.cpp")
    deps = [src, os.path.join(ROOT, "naf_b200/csrc/zstd_dec.cuh"), os.path.join(ROOT, "naf_b200/csrc/zstd_hd.cuh")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-Wall", "-Wno-unused-function", "-o", EXE, src], check=True)
    return EXE


def test_emu_decoder_on_golden_frames(emu, tmp_path):
    out = str(tmp_path / "o.bin")
    for e in helpers.manifest("zstd"):
        p = subprocess.run([emu, os.path.join(helpers.GOLDEN, "zstd", e["frame"]), out], capture_output=True)
        assert p.returncode == 0, (e["frame"], p.stderr)
        assert helpers.sha(open(out, "rb").read()) == e["sha256"], e["frame"]


@pytest.mark.skipif(not os.access(os.path.join(helpers.REF_BIN, "decodecorpus"), os.X_OK), reason="oracle/_ref/decodecorpus not built")
def test_fresh_decodecorpus_frames(emu, tmp_path):
    """frames the committed goldens have not seen: zstd's own generator of valid frames (every block / literal / sequence mode,
    repeat offsets, treeless literals, Repeat_Mode tables) with a different seed, including the large ones the goldens leave out
    for size; the oracle's from-spec decoder and our decoder's HD bodies must both regenerate the originals"""
    dc, dco = tmp_path / "dc", tmp_path / "dco"
    dc.mkdir(); dco.mkdir()
    seed = int(os.environ.get("NAF_DC_SEED", "77"))
    subprocess.run([os.path.join(helpers.REF_BIN, "decodecorpus"), "-n120", f"-s{seed}", f"-p{dc}", f"-o{dco}"], capture_output=True, check=True)
    oracle = helpers.load_oracle()
    out, n = str(tmp_path / "o.bin"), 0
    for f in sorted(os.listdir(dc)):
        z = open(dc / f, "rb").read()
        raw = open(dco / f[:-4], "rb").read()
        if len(raw) > (8 << 20):
            continue
        assert oracle.zstd_decompress(z) == raw, f
        p = subprocess.run([emu, str(dc / f), out], capture_output=True)
        assert p.returncode == 0, (f, p.stderr)
        assert open(out, "rb").read() == raw, f
        n += 1
    assert n > 60
