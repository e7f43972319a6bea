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


def test_damaged_frames_with_inherited_tables_never_leave_the_buffers(tmp_path):
    """bit flips, truncation and overwritten spans in frames whose blocks inherit their tables (Treeless_Literals / Repeat_Mode: what
    the level >= 2 encoder writes, csrc/zstd_lzc_hd.cuh): the decoder's HD bodies, built with the address and undefined-behaviour
    sanitizers, either report an error or regenerate some bytes (no checksum) -- they never read or write outside their buffers.
    The GPU counterpart (`test_corrupt_input_never_crashes`) can only see that the context survives."""
    import random
    import struct
    src = os.path.join(ROOT, "tests", "emu", "emu_zstd.cpp")
    exe = os.path.join(ROOT, "tests", "_build", "emu_zstd_asan")
    deps = [src, os.path.join(ROOT, "naf_b200/csrc/zstd_dec.cuh"), os.path.join(ROOT, "naf_b200/csrc/zstd_hd.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        p = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", exe, src], capture_output=True)
        if p.returncode:
            pytest.skip("no sanitizer runtime on this machine")
    enc = os.path.join(ROOT, "tests", "_build", "emu_zlzc_bytes")
    esrc = os.path.join(ROOT, "tests", "emu", "emu_zlzc.cpp")
    if not os.path.exists(enc):
        subprocess.run(["g++", "-std=c++17", "-O2", "-DZLC_BYTES", "-o", enc, esrc], check=True)
    rng = random.Random(7)
    frames = []
    for data in (b"".join(b"SRR1.%d\0" % i for i in range(1, 6000)), b"".join(b"%d/1 length=%d\0" % (i, 100 + i % 50) for i in range(1, 4000)),
                 struct.pack("<I", 150) * 9000, b"x" * 9000 + b"".join(b"r%d\0" % (i % 9) for i in range(9000))):
        inp, z = str(tmp_path / "i.bin"), str(tmp_path / "c.zst")
        with open(inp, "wb") as f:
            f.write(data)
        assert subprocess.run([enc, inp, z, "8192"], capture_output=True).returncode == 0
        frames.append(open(z, "rb").read())
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=1")
    errors = 0
    for it in range(120):
        fr = bytearray(rng.choice(frames))
        kind = rng.random()
        if kind < 0.5:
            for _ in range(rng.randint(1, 3)):
                fr[rng.randrange(len(fr))] ^= 1 << rng.randrange(8)
        elif kind < 0.75:
            fr = fr[:rng.randrange(1, len(fr))]
        else:
            at = rng.randrange(len(fr))
            fr[at:at + rng.randint(1, 64)] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 64)))
        z, out = str(tmp_path / "d.zst"), str(tmp_path / "d.out")
        with open(z, "wb") as f:
            f.write(fr)
        p = subprocess.run([exe, z, out], capture_output=True, env=env, timeout=120)
        assert p.returncode in (0, 1) and b"Sanitizer" not in p.stderr and b"runtime error" not in p.stderr, (it, p.returncode, p.stderr[-800:])
        errors += p.returncode == 1
    assert errors > 40
