"""not-gpu: the block encoder's HD bodies (naf_b200/csrc/zstd_enc_hd.cuh: LZ77 match finder, Huffman literals, FSE-coded
sequences, repeat-offset codes) run on the CPU by tests/emu/emu_zenc.cpp.  Every frame must be decoded back to the input
by (a) the oracle's from-spec decoder, (b) the unmodified libzstd 1.5.0 built from the reference's sources (when
present), (c) our own decoder's HD bodies (tests/emu/emu_zstd).  The encoder is free-parse, so parity on this side IS
decodability; the ratio checks only guard against the match finder silently finding nothing."""
import ctypes as C
import os
import random
import struct
import subprocess

import numpy as np
import pytest

import helpers
from naf_b200 import synth

ROOT = helpers.ROOT
BUILD = os.path.join(ROOT, "tests", "_build")
CSRC = os.path.join(ROOT, "naf_b200", "csrc")


def _build(name, deps, source=None, flags=()):
    os.makedirs(BUILD, exist_ok=True)
    exe, src = os.path.join(BUILD, name), os.path.join(ROOT, "tests", "emu", (source or name) + ".cpp")
    deps = [src, os.path.join(ROOT, "tests", "emu", "lzcol.hpp")] + [os.path.join(CSRC, d) for d in deps]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-Wall", "-Wno-unused-function", *flags, "-o", exe, src], check=True)
    return exe


ZLC_DEPS = ["zstd_lzc_hd.cuh", "zstd_lzc_bytes_hd.cuh", "zstd_enc_hd.cuh", "zstd_hd.cuh"]


def _build_zlzc():
    """tests/emu/emu_zlzc.cpp with the finder's phases as byte loops (what a level >= 2 runs on the GPU) and as bit masks (NAFGPU_LZ=b)"""
    return [_build("emu_zlzc_bytes", ZLC_DEPS, source="emu_zlzc", flags=("-DZLC_BYTES",)), _build("emu_zlzc", ZLC_DEPS)]


@pytest.fixture(scope="module")
def enc():
    return _build("emu_zenc", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])


@pytest.fixture(scope="module")
def dec():
    return _build("emu_zstd", ["zstd_dec.cuh", "zstd_hd.cuh"])


@pytest.fixture(scope="module")
def libzstd():
    so = os.path.join(helpers.REF_BIN, "libzstd.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    lib.ZSTD_decompress.restype = C.c_size_t
    lib.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.ZSTD_isError.argtypes = [C.c_size_t]
    # frames without a content size: streaming API
    lib.ZSTD_createDStream.restype = C.c_void_p
    lib.ZSTD_freeDStream.argtypes = [C.c_void_p]
    lib.ZSTD_decompressStream.restype = C.c_size_t
    lib.ZSTD_decompressStream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ZSTD_getErrorName.restype = C.c_char_p
    lib.ZSTD_getErrorName.argtypes = [C.c_size_t]
    return lib


class _Buf(C.Structure):
    _fields_ = [("p", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


def libzstd_decode(lib, frame, expect):
    ds = lib.ZSTD_createDStream()
    src = C.create_string_buffer(frame, len(frame))
    dst = C.create_string_buffer(expect + 64)
    i, o = _Buf(C.cast(src, C.c_void_p), len(frame), 0), _Buf(C.cast(dst, C.c_void_p), expect + 64, 0)
    while True:
        r = lib.ZSTD_decompressStream(ds, C.byref(o), C.byref(i))
        assert not lib.ZSTD_isError(r), lib.ZSTD_getErrorName(r)
        if r == 0 or i.pos == i.size:
            break
    lib.ZSTD_freeDStream(ds)
    assert r == 0, "libzstd wants more input"
    return dst.raw[:o.pos]


def roundtrip(enc, dec, libzstd, tmp_path, data, block=8192, lz=1, hstride=1, own_decoder=True):
    inp, z, back = str(tmp_path / "in.bin"), str(tmp_path / "f.zst"), str(tmp_path / "back.bin")
    with open(inp, "wb") as f:
        f.write(data)
    p = subprocess.run([enc, inp, z, str(block), str(lz), str(hstride)], capture_output=True)
    assert p.returncode == 0, p.stderr
    frame = open(z, "rb").read()
    assert helpers.load_oracle().zstd_decompress(frame) == data, "oracle decoder"
    if libzstd is not None:
        assert libzstd_decode(libzstd, frame, len(data)) == data, "libzstd 1.5.0"
    if own_decoder:
        q = subprocess.run([dec, z, back], capture_output=True)
        assert q.returncode == 0, q.stderr
        assert open(back, "rb").read() == data, "own decoder (HD bodies)"
    return len(frame)


def ids_stream(n, start=1):
    return b"".join(b"SRR1.%d\0" % i for i in range(start, start + n))


def test_ids_comments_lengths(enc, dec, libzstd, tmp_path):
    ids = ids_stream(20000, 999000)
    z = roundtrip(enc, dec, libzstd, tmp_path, ids)
    assert z < 0.2 * len(ids), (z, len(ids))           # Huffman-only gets ~0.45; zstd -1 ~0.1
    comm = b"".join(b"%d/1\0" % i for i in range(1, 20001))
    assert roundtrip(enc, dec, libzstd, tmp_path, comm) < 0.25 * len(comm)
    lens = struct.pack("<I", 150) * 50000
    assert roundtrip(enc, dec, libzstd, tmp_path, lens) < 0.005 * len(lens)
    illumina = b"".join(b"A00123:45:HXXXX:1:%d:%d:%d\0" % (1101 + i // 500, 1000 + (i * 37) % 9000, 2000 + (i * 91) % 30000) for i in range(8000))
    assert roundtrip(enc, dec, libzstd, tmp_path, illumina) < 0.45 * len(illumina)
    ont = b"".join(b"ont_%d len=%d\0" % (i, 10000 + (i * 7919) % 40000) for i in range(10000))
    roundtrip(enc, dec, libzstd, tmp_path, ont)


@pytest.mark.parametrize("block", [8192, 32768, 1024, 100])
def test_block_sizes_and_modes(enc, dec, libzstd, tmp_path, block):
    rng = np.random.default_rng(block)
    cases = [
        b"", b"a", b"ab" * 7, b"x" * 15, b"x" * 16, b"x" * 100000,                       # empty / raw / RLE blocks
        bytes(rng.integers(0, 256, 50000, dtype=np.uint8)),                             # incompressible: raw blocks
        bytes(rng.integers(0, 4, 70000, dtype=np.uint8)),                               # Huffman, few matches
        bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 40001)),
        b"abcdefgh" * 5000 + b"tail",                                                   # one long periodic match per block
        b"0123456789" * 3 + bytes(rng.integers(0, 256, 30, dtype=np.uint8)) + b"0123456789" * 900,
        ids_stream(3000) + bytes(rng.integers(0, 256, 3000, dtype=np.uint8)) + ids_stream(3000, 5),
        b"".join(bytes([65 + (i * i) % 23]) * (1 + i % 40) for i in range(3000)),        # runs: offset-1 matches, ll == 0 cases
        b"".join(struct.pack("<I", int(v)) for v in rng.integers(10000, 50000, 9000)),  # ONT-like length units
        bytes([255]) * 30000 + bytes(rng.integers(0, 255, 5000, dtype=np.uint8)),       # mask units
    ]
    for data in cases:
        roundtrip(enc, dec, libzstd, tmp_path, data, block=block)
        roundtrip(enc, dec, libzstd, tmp_path, data, block=block, lz=0, own_decoder=False)   # literals-only path of the same coder


def test_interleaved_hash_table_is_the_same_parse(enc, dec, libzstd, tmp_path):
    """the kernel interleaves the hash tables of a warp's lanes (stride 32): same frame as stride 1"""
    data = ids_stream(5000, 12345)
    inp = str(tmp_path / "i.bin")
    open(inp, "wb").write(data)
    outs = []
    for stride in (1, 32):
        z = str(tmp_path / f"s{stride}.zst")
        assert subprocess.run([enc, inp, z, "8192", "1", str(stride)], capture_output=True).returncode == 0
        outs.append(open(z, "rb").read())
    assert outs[0] == outs[1]


def test_fuzz_structured(enc, dec, libzstd, tmp_path):
    """random mixtures of copies, runs and noise: every repeat-offset / literal-length-0 / mode combination gets hit"""
    rnd = random.Random(7)
    for it in range(60):
        n = rnd.choice([17, 200, 5000, 8192, 8193, 20000, 40000])
        out = bytearray()
        alpha = rnd.choice([2, 4, 16, 64, 256])
        while len(out) < n:
            k = rnd.random()
            if k < 0.45 and len(out) > 8:
                off = rnd.choice([1, 2, 3, 4, 13, rnd.randint(1, len(out))])
                off = min(off, len(out))
                ln = rnd.choice([3, 4, 5, 8, 20, 300, rnd.randint(3, 2000)])
                for _ in range(ln):
                    out.append(out[-off])
            elif k < 0.55:
                out += bytes([rnd.randrange(alpha)]) * rnd.randint(1, 50)
            else:
                out += bytes(rnd.randrange(alpha) for _ in range(rnd.randint(1, 30)))
        data = bytes(out[:n])
        roundtrip(enc, dec, libzstd, tmp_path, data, block=rnd.choice([8192, 8192, 4096, 32768, 333]))


def test_golden_texts(enc, dec, libzstd, tmp_path):
    """the six streams of real-shaped inputs (oracle split of synthetic FASTQ / soft-masked FASTA)"""
    oracle = helpers.load_oracle()
    for text in (synth.fastq(3000, 150, seed=3), synth.ont_fasta(40, 1000, 5000, seed=4), synth.protein_fasta(500, 300, seed=5)):
        kw = {"seq_type": "protein"} if text.startswith(b">sp|") else {}
        streams, _ = oracle.split(text, **kw)
        for s in streams:
            roundtrip(enc, dec, libzstd, tmp_path, s)


def test_warp_match_finder_prototype(libzstd, tmp_path):
    """tests/emu/proto_lzw.cpp: the window-at-a-time match finder planned for the next round (what a warp would run), feeding
    the shipped literal / sequence coder.  Valid frames, sizes no worse than the serial parse, and about one step per sequence."""
    exe = _build("proto_lzw", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    enc = _build("emu_zenc", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    rng = np.random.default_rng(5)
    for data in (ids_stream(20000, 999000), b"".join(b"%d/1\0" % i for i in range(1, 20001)), struct.pack("<I", 150) * 50000,
                 bytes(rng.integers(0, 4, 70000, dtype=np.uint8)), b"".join(bytes([65 + (i * i) % 23]) * (1 + i % 40) for i in range(3000)),
                 b"", b"x" * 9000, bytes(rng.integers(0, 256, 20000, dtype=np.uint8))):
        inp, z, zs = str(tmp_path / "i.bin"), str(tmp_path / "w.zst"), str(tmp_path / "s.zst")
        with open(inp, "wb") as f:
            f.write(data)
        p = subprocess.run([exe, inp, z, "8192"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        frame = open(z, "rb").read()
        assert helpers.load_oracle().zstd_decompress(frame) == data
        if libzstd is not None:
            assert libzstd_decode(libzstd, frame, len(data)) == data
        assert subprocess.run([enc, inp, zs, "8192", "1"], capture_output=True).returncode == 0
        assert len(frame) <= os.path.getsize(zs) * 1.02 + 16
        stats = dict(kv.split("=") for kv in p.stdout.split())
        assert int(stats["steps"]) <= int(stats["seqs"]) + int(stats["blocks"]) * 257


def test_column_match_finder_prototype(libzstd, tmp_path):
    """tests/emu/proto_lzcol.cpp: a match finder made of maps and scans only (candidate offset = the same column of the previous
    '\\0'-terminated record, else the previous 4-byte unit; runs of matching bytes are the matches), feeding the shipped literal /
    sequence coder.  Valid frames on anything; on the streams it is for, within 15 % of the serial hash parse."""
    exe = _build("proto_lzcol", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    enc = _build("emu_zenc", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    rng = np.random.default_rng(6)
    text_like = (ids_stream(20000, 999000), b"".join(b"%d/1\0" % i for i in range(1, 20001)), struct.pack("<I", 150) * 50000)
    other = (b"".join(struct.pack("<I", 100 + (i * 7) % 3) for i in range(30000)),       # period of three units: not a candidate it has
             bytes(rng.integers(0, 4, 70000, dtype=np.uint8)), b"".join(bytes([65 + (i * i) % 23]) * (1 + i % 40) for i in range(3000)),
             b"", b"x" * 9000, bytes(rng.integers(0, 256, 20000, dtype=np.uint8)), b"\0" * 5000 + b"ab\0" * 3000, b"no terminator at all " * 700)
    for data in text_like + other:
        for bs in ("8192", "1000"):
            inp, z, zs = str(tmp_path / "i.bin"), str(tmp_path / "c.zst"), str(tmp_path / "s.zst")
            with open(inp, "wb") as f:
                f.write(data)
            p = subprocess.run([exe, inp, z, bs], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            frame = open(z, "rb").read()
            assert helpers.load_oracle().zstd_decompress(frame) == data
            if libzstd is not None:
                assert libzstd_decode(libzstd, frame, len(data)) == data
            if data in text_like and bs == "8192":
                assert subprocess.run([enc, inp, zs, bs, "1"], capture_output=True).returncode == 0
                assert len(frame) <= os.path.getsize(zs) * 1.15 + 64, (len(frame), os.path.getsize(zs))


def test_shared_table_block_coder_prototype(libzstd, tmp_path):
    """tests/emu/proto_shared.cpp: one Huffman code and one set of FSE tables per STREAM (from a sample of its blocks); the first
    block carries them, every later block is Treeless_Literals + Repeat_Mode -- what a block then costs is serial coding against
    read-only tables.  libzstd and the oracle must decode the frames (they track the tables across blocks, raw and RLE blocks in
    between included); sizes stay within 1.4 x of per-block tables on the text-like streams."""
    exe = _build("proto_shared", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])      # (also includes tests/emu/lzcol.hpp)
    enc = _build("emu_zenc", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    rng = np.random.default_rng(8)
    text_like = (ids_stream(60000, 1), b"".join(b"%d/1\0" % i for i in range(1, 60001)), struct.pack("<I", 150) * 50000)
    other = (b"x" * 9000 + ids_stream(3000, 5) + b"y" * 20000 + ids_stream(3000, 77777),          # RLE blocks before and between
             bytes(rng.integers(0, 256, 9000, dtype=np.uint8)) + ids_stream(4000, 123),           # a raw block first
             ids_stream(3000, 1) + bytes(rng.integers(0, 256, 20000, dtype=np.uint8)) + ids_stream(3000, 9),   # bytes the sample never saw
             bytes(rng.integers(0, 4, 70000, dtype=np.uint8)), b"", b"q" * 30000, bytes(rng.integers(0, 256, 20000, dtype=np.uint8)),
             b"".join(bytes([65 + (i * i) % 23]) * (1 + i % 40) for i in range(3000)))
    for data in text_like + other:
        for bs in ("8192", "2048", "1000"):
            inp, z, zs = str(tmp_path / "i.bin"), str(tmp_path / "c.zst"), str(tmp_path / "s.zst")
            with open(inp, "wb") as f:
                f.write(data)
            for finder in ([], ["col"]):                      # sequences from the serial hash parse / from proto_lzcol's column finder
                p = subprocess.run([exe, inp, z, bs] + finder, capture_output=True, text=True)
                assert p.returncode == 0, p.stderr
                frame = open(z, "rb").read()
                assert helpers.load_oracle().zstd_decompress(frame) == data, (len(data), bs, finder)
                if libzstd is not None:
                    assert libzstd_decode(libzstd, frame, len(data)) == data, (len(data), bs, finder)
                if data in text_like and bs == "8192":
                    assert subprocess.run([enc, inp, zs, bs, "1"], capture_output=True).returncode == 0
                    assert len(frame) <= os.path.getsize(zs) * 1.4 + 64, (len(frame), os.path.getsize(zs), finder)


def _zlzc_cases():
    rng = np.random.default_rng(9)
    rnd = random.Random(9)
    text_like = [ids_stream(60000, 1), b"".join(b"%d/1\0" % i for i in range(1, 60001)), struct.pack("<I", 150) * 50000,
                 b"".join(b"A00123:45:HXXXX:1:%d:%d:%d\0" % (1101 + i // 500, 1000 + (i * 37) % 9000, 2000 + (i * 91) % 30000) for i in range(8000))]
    other = [b"x" * 9000 + ids_stream(3000, 5) + b"y" * 20000 + ids_stream(3000, 77777),          # RLE blocks before and between
             bytes(rng.integers(0, 256, 9000, dtype=np.uint8)) + ids_stream(4000, 123),           # a raw block first
             ids_stream(3000, 1) + bytes(rng.integers(0, 256, 20000, dtype=np.uint8)) + ids_stream(3000, 9),   # bytes the sample never saw
             bytes(rng.integers(0, 4, 70000, dtype=np.uint8)), b"", b"a", b"q" * 30000, bytes(rng.integers(0, 256, 20000, dtype=np.uint8)),
             b"".join(bytes([65 + (i * i) % 23]) * (1 + i % 40) for i in range(3000)),            # runs: candidate 4 against the column
             b"".join(struct.pack("<I", int(v)) for v in rng.integers(10000, 50000, 9000)),      # ONT-like length units: no sequences at all
             bytes([255]) * 30000 + bytes(rng.integers(0, 255, 5000, dtype=np.uint8)),           # mask units
             b"\0" * 5000 + b"ab\0" * 3000 + b"\0\0abcdefgh\0" * 500, b"no terminator at all " * 700,
             b"".join(b"r%d\0" % (i % 7) * (1 + i % 3) for i in range(9000)),                     # records of changing length, runs across records
             ids_stream(700, 1)[:8192] + b"z" * 8192 + ids_stream(5000, 4)]                       # exactly one full block, then RLE, then more
    for _ in range(12):                                                                           # mixtures of copies, runs, noise and terminators
        n, out, alpha = rnd.choice([17, 200, 5000, 8192, 8193, 20000, 40000]), bytearray(), rnd.choice([2, 4, 16, 64, 256])
        while len(out) < n:
            k = rnd.random()
            if k < 0.45 and len(out) > 8:
                off = min(rnd.choice([1, 2, 3, 4, 13, rnd.randint(1, len(out))]), len(out))
                for _ in range(rnd.choice([3, 4, 5, 8, 20, 300, rnd.randint(3, 2000)])):
                    out.append(out[-off])
            elif k < 0.55:
                out += bytes([rnd.randrange(alpha)]) * rnd.randint(1, 50)
            elif k < 0.65:
                out += b"\0"
            else:
                out += bytes(rnd.randrange(alpha) for _ in range(rnd.randint(1, 30)))
        other.append(bytes(out[:n]))
    # the inputs tests/test_gpu_encode.py compresses at level 3 on the device (the GPU writes what the emulation writes)
    r3 = np.random.default_rng(3)
    other += [b"A", b"AB", b"A" * 70000, bytes(range(256)) * 300]
    for n in [2, 3, 15, 16, 17, 255, 256, 1023, 1024, 1025, 5000, 32767, 32768, 32769, 65535, 65536, 65537, 200000]:
        other.append(bytes(r3.choice(np.frombuffer(b"\x11\x12\x14\x18\x21\x22\x24\x28\x41\x42\x44\x48\x81\x82\x84\x88", dtype=np.uint8), n)))
        other.append((np.clip(np.round(r3.normal(34, 6, n)), 2, 40).astype(np.uint8) + 33).tobytes())
    other += [r3.integers(0, 256, 300000, dtype=np.uint8).tobytes(),
              bytes(r3.choice(np.arange(200, dtype=np.uint8), 100000, p=np.r_[[0.5], np.full(199, 0.5 / 199)])),
              b"abcdefgh" * 9000 + b"tail", b"r7\0" * 5000 + bytes(r3.integers(0, 256, 9000, dtype=np.uint8))]
    return text_like, other


def test_column_finder_and_stream_tables(dec, libzstd, tmp_path):
    """naf_b200/csrc/zstd_lzc_hd.cuh -- the data-parallel LZ stage (NAFGPU_LZ=shared): the column match finder as the phases one
    CTA runs (thread = 32-byte chunk), per-stream Huffman / FSE tables from sampled blocks, blocks coded against them
    (Treeless_Literals + Repeat_Mode).  tests/emu/emu_zlzc.cpp runs the phases thread by thread.  Every frame is decoded by the
    oracle, libzstd 1.5.0 and our own decoder, and equals byte for byte what the serial restatement (lzcol.hpp + proto_shared.cpp)
    writes; on the streams it is for the result stays within 1.4 x of per-block tables + the serial hash parse."""
    exe, exe_bits = _build_zlzc()
    proto = _build("proto_shared", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    enc = _build("emu_zenc", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    oracle = helpers.load_oracle()
    text_like, other = _zlzc_cases()
    streams, _ = oracle.split(synth.fastq(3000, 150, seed=3))
    other += [s for s in streams[:4]]
    streams, _ = oracle.split(synth.ont_fasta(40, 1000, 5000, seed=4))
    other += [s for s in streams[:4]]
    for idx, data in enumerate(text_like + other):
        for bs in ("8192", "2048", "1000", "100", "33") if idx % 3 == 0 or len(data) < 30000 else ("8192", "1000"):
            inp, z, zp, zs, back = (str(tmp_path / x) for x in ("i.bin", "c.zst", "p.zst", "s.zst", "back.bin"))
            with open(inp, "wb") as f:
                f.write(data)
            p = subprocess.run([exe, inp, z, bs], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            frame = open(z, "rb").read()
            assert subprocess.run([exe_bits, inp, zp, bs], capture_output=True).returncode == 0
            assert open(zp, "rb").read() == frame, ("the two formulations of the finder differ", len(data), bs)
            assert oracle.zstd_decompress(frame) == data, (len(data), bs)
            if libzstd is not None:
                assert libzstd_decode(libzstd, frame, len(data)) == data, (len(data), bs)
            q = subprocess.run([dec, z, back], capture_output=True)
            assert q.returncode == 0 and open(back, "rb").read() == data, (len(data), bs, q.stderr)
            if int(bs) >= 64:
                assert subprocess.run([proto, inp, zp, bs, "col"], capture_output=True).returncode == 0
                assert open(zp, "rb").read() == frame, ("serial restatement differs", len(data), bs)
            if any(data is t for t in text_like) and bs == "8192":
                assert subprocess.run([enc, inp, zs, bs, "1"], capture_output=True).returncode == 0
                assert len(frame) <= os.path.getsize(zs) * 1.4 + 64, (len(frame), os.path.getsize(zs))


def _frame_blocks(frame):
    """-> the block section of a single zstd frame written by our encoders (FHD 0x00 + window byte), as a list of (header int, content)"""
    assert frame[:4] == b"\x28\xb5\x2f\xfd" and frame[4] == 0
    at, out = 6, []
    while True:
        bh = frame[at] | (frame[at + 1] << 8) | (frame[at + 2] << 16)
        size = 1 if (bh >> 1) & 3 == 1 else bh >> 3
        out.append((bh, frame[at + 3:at + 3 + size]))
        at += 3 + size
        if bh & 1:
            break
    assert at == len(frame)
    return out


def test_stream_tables_survive_shard_concatenation(libzstd, tmp_path):
    """N GPUs encode consecutive shards of a stream and the blocks are concatenated into ONE frame (sharded.py /
    nafgpu_shard_finish).  With per-stream tables every shard's first coded block carries that shard's tables and the blocks behind
    it inherit them (Treeless_Literals / Repeat_Mode) -- so the merged frame must decode whatever the neighbours are: shards with
    tables, shards that fell back to per-block tables, raw / RLE-only shards, empty shards."""
    exe = _build_zlzc()[0]
    oracle = helpers.load_oracle()
    rng = np.random.default_rng(4)
    ids = ids_stream(9000, 1)
    parts_sets = [
        [ids[:30000], ids[30000:61000], ids[61000:]],                                                   # three shards with tables of their own
        [ids[:20000], bytes(rng.integers(0, 256, 20000, dtype=np.uint8)), ids[20000:50000]],            # raw-only shard in the middle
        [b"x" * 20000, ids[:25000], b"", struct.pack("<I", 150) * 6000, ids[25000:40000]],               # RLE-only, empty, other statistics
        [b"".join(struct.pack("<I", int(v)) for v in rng.integers(10000, 50000, 5000)), ids[:30000]],    # per-block tables first, then shared
        [ids[:30000], b"".join(struct.pack("<I", int(v)) for v in rng.integers(10000, 50000, 5000)), ids[30000:45000]],
    ]
    for parts in parts_sets:
        merged = bytearray(b"\x28\xb5\x2f\xfd\x00" + bytes([(17 - 10) << 3]))
        for k, part in enumerate(parts):
            inp, z = str(tmp_path / "i.bin"), str(tmp_path / "c.zst")
            with open(inp, "wb") as f:
                f.write(part)
            assert subprocess.run([exe, inp, z, "8192"], capture_output=True).returncode == 0
            blocks = _frame_blocks(open(z, "rb").read())
            for j, (bh, content) in enumerate(blocks):
                last = k == len(parts) - 1 and j == len(blocks) - 1
                bh = (bh & ~1) | int(last)
                merged += bytes([bh & 0xFF, (bh >> 8) & 0xFF, bh >> 16]) + content
        data = b"".join(parts)
        assert oracle.zstd_decompress(bytes(merged)) == data
        if libzstd is not None:
            assert libzstd_decode(libzstd, bytes(merged), len(data)) == data


def test_defining_block_that_does_not_fit_gets_the_stream_private_tables(dec, libzstd, tmp_path):
    """the first block with literals and sequences is mostly noise, the sample is dominated by 860 blocks of names: coded with the
    stream's Huffman code the noise overflows the block's slot, the defining block cannot be written, and the whole stream falls
    back to per-block tables (k_zlc_finish_own) -- still byte for byte the serial restatement, still valid for every decoder"""
    exe, exe_bits = _build_zlzc()
    proto = _build("proto_shared", ["zstd_enc_hd.cuh", "zstd_hd.cuh"])
    rng = np.random.default_rng(3)
    b0 = bytearray(rng.integers(1, 256, 8192, dtype=np.uint8).tobytes())
    for i in list(range(100, 130)) + list(range(5000, 5040)):
        b0[i] = b0[i - 4]
    data = bytes(b0) + ids_stream(600000, 1)
    inp, z, zp, back = (str(tmp_path / x) for x in ("i.bin", "c.zst", "p.zst", "back.bin"))
    with open(inp, "wb") as f:
        f.write(data)
    p = subprocess.run([exe, inp, z, "8192"], capture_output=True, text=True)
    assert p.returncode == 0 and "def_fail=1" in p.stdout, p.stdout
    frame = open(z, "rb").read()
    assert helpers.load_oracle().zstd_decompress(frame) == data
    if libzstd is not None:
        assert libzstd_decode(libzstd, frame, len(data)) == data
    assert subprocess.run([dec, z, back], capture_output=True).returncode == 0 and open(back, "rb").read() == data
    assert subprocess.run([proto, inp, zp, "8192", "col"], capture_output=True).returncode == 0
    assert open(zp, "rb").read() == frame
    assert subprocess.run([exe_bits, inp, zp, "8192"], capture_output=True).returncode == 0
    assert open(zp, "rb").read() == frame


def test_reference_unnaf_decodes_files_with_stream_table_frames(tmp_path):
    """a whole .naf whose ids / comments / lengths sections are the frames the data-parallel stage writes (emulation; the GPU writes
    the same bytes), the other sections as the oracle's encoder makes them: the UNMODIFIED reference unnaf (input.c:211
    ZSTD_decompress for these sections) prints the input back, and so does the oracle"""
    from naf_b200 import container
    exe = _build_zlzc()[0]
    oracle = helpers.load_oracle()
    for text, kw in [(synth.fastq(30_000, 150, seed=21), {}), (synth.ont_fasta(60, 10000, 30000, seed=22), {}),
                     (synth.protein_fasta(8000, 300, seed=23), {"seq_type": "protein"})]:
        naf, _ = oracle.encode(text, **kw)
        streams, info = oracle.split(text, **kw)
        h = container.read_header(naf)
        first = min(sec[2] for sec in h.sections if sec is not None)
        out = bytearray()
        for k in range(6):
            sec = h.sections[k]
            if sec is None:
                continue
            orig, comp, off = sec
            body = naf[off:off + comp]
            if k < 3:
                inp, z = str(tmp_path / "i.bin"), str(tmp_path / "c.zst")
                with open(inp, "wb") as f:
                    f.write(streams[k])
                assert subprocess.run([exe, inp, z, "8192"], capture_output=True).returncode == 0
                body = open(z, "rb").read()[4:]                # sections are stored without the 4-byte magic (compressor.c:158)
                assert orig == len(streams[k])
            out += container.put_vle(orig) + container.put_vle(len(body)) + body
        # header bytes up to the first section's two VLE numbers
        hdr_end = first - len(container.put_vle(h.sections[0][0])) - len(container.put_vle(h.sections[0][1]))
        mixed = bytes(naf[:hdr_end]) + bytes(out)
        assert oracle.decode(mixed) == text
        if helpers.have_ref():
            rc, got, err = helpers.ref_run("unnaf", [], mixed)
            assert rc == 0 and got == text, err[:300]


def test_finder_phases_do_not_depend_on_thread_order(tmp_path):
    """between two barriers the threads of a CTA run in no particular order: the emulation shuffles the order of the threads of every
    phase (ZLC_ORDER=<seed>); a phase in which one thread read what another one writes would give different frames.  Both
    formulations of the finder, several block sizes."""
    exes = _build_zlzc()
    text_like, other = _zlzc_cases()
    picks = text_like[:3] + [d for d in other if 0 < len(d) <= 100000][:14]
    for data in picks:
        inp = str(tmp_path / "i.bin")
        with open(inp, "wb") as f:
            f.write(data[:200000])
        for bs in ("8192", "1000", "100"):
            for exe in exes:
                frames = []
                for seed in ("0", "1", "12345"):
                    z = str(tmp_path / ("o%s.zst" % seed))
                    assert subprocess.run([exe, inp, z, bs], capture_output=True, env=dict(os.environ, ZLC_ORDER=seed)).returncode == 0
                    frames.append(open(z, "rb").read())
                assert frames[0] == frames[1] == frames[2], (len(data), bs, os.path.basename(exe))
