"""One .naf from several shards (naf_b200/sharded.py).

not-gpu: the exchange / link / layout logic, with a CPU shard encoder built on the oracle (raw zstd blocks): in one
process for many shard counts and edge cases, and over a real world_size-2 gloo process group.
gpu: the same protocol driving libnafgpu.so (nafgpu_shard_begin / _finish / _fetch), several contexts on one GPU.
In every case the merged file must decode — by the oracle and, where built, by the UNMODIFIED reference unnaf — to
exactly the concatenation of the shards' texts."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers
from naf_b200 import api, sharded, synth

CODE_OF = np.full(256, 15, dtype=np.uint8)
for _i, _c in enumerate(b"-TGKCYSBAWRDMHVN"):
    CODE_OF[_c] = _i
    CODE_OF[_c | 0x20] = _i
CODE_OF[ord("U")] = CODE_OF[ord("u")] = 1


class OracleShardEncoder:
    """CPU stand-in for GpuShardEncoder: oracle.split for the shard's streams, numpy for the link step, raw zstd blocks."""

    def __init__(self, oracle, **kw):
        self.oracle, self.kw = oracle, kw

    def begin(self, text, opts):
        streams, info = self.oracle.split(bytes(text), **self.kw)
        self.streams, self.info = streams, info
        self.packed = self.kw.get("seq_type", "dna") in ("dna", "rna")
        n = info["seq_size"]
        c = sharded.Counts(n_records=info["n_sequences"], n_bases=n, longest_line=info["longest_line"], format=info["format"])
        if self.packed:
            p = np.frombuffer(streams[4], dtype=np.uint8)
            codes = np.empty(2 * len(p), dtype=np.uint8)
            codes[0::2], codes[1::2] = p & 15, p >> 4
            self.codes = codes[:n]
            case = np.zeros(n, dtype=np.uint8)
            if info["store_mask"]:
                pos, on, run = 0, 0, 0
                for u in streams[3]:
                    run += u
                    if u != 255:
                        case[pos:pos + run] = on
                        pos, run, on = pos + run, 0, on ^ 1
            self.case = case
            if n:
                flips = np.flatnonzero(case[1:] != case[:-1]) + 1
                c.first_code, c.first_case, c.last_case = int(self.codes[0]), int(case[0]), int(case[-1])
                c.n_flips, c.last_flip = len(flips), int(flips[-1]) if len(flips) else 0
        return c

    def finish(self, link):
        s = list(self.streams)
        if self.packed:
            codes = self.codes[1:] if (link.bases_before & 1) and len(self.codes) else self.codes
            if len(codes) & 1:
                codes = np.append(codes, np.uint8(link.next_first_code))
            s[4] = (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8).tobytes()
            if self.info["store_mask"]:
                n = len(self.case)
                flips = list(np.flatnonzero(self.case[1:] != self.case[:-1]) + 1) if n else []
                if n and int(self.case[0]) != link.prev_last_case:
                    flips = [0] + flips
                units, prev = bytearray(), -link.run_carry
                for f in flips:
                    L = int(f) - prev
                    units += b"\xff" * (L // 255) + bytes([L % 255])
                    prev = int(f)
                if link.is_last and n - prev > 0:
                    L = n - prev
                    units += b"\xff" * (L // 255) + bytes([L % 255])
                s[3] = bytes(units)
            else:
                s[3] = b""
        if not self.info["store_qual"]:
            s[5] = b""
        self.bodies = []
        for k in range(6):
            data, body = s[k], bytearray()
            nblk = max(1, (len(data) + 131071) // 131072)
            for b in range(nblk):
                piece = data[b * 131072:(b + 1) * 131072]
                last = 1 if (link.is_last and b == nblk - 1) else 0
                h = last | (0 << 1) | (len(piece) << 3)
                body += bytes([h & 255, (h >> 8) & 255, (h >> 16) & 255]) + piece
            self.bodies.append(bytes(body))
        present = [True, True, True, self.packed and not self.kw.get("no_mask"), True, bool(self.info["store_qual"]) or bool(link.store_qual)]
        raw = [len(s[k]) if present[k] else 0 for k in range(6)]
        return raw, [len(self.bodies[k]) if present[k] else 0 for k in range(6)]

    def fetch(self, k, dst):
        import torch
        if dst.numel():
            dst.copy_(torch.frombuffer(bytearray(self.bodies[k]), dtype=torch.uint8))


def _opts(**kw):
    return api.make_enc_opts(**kw)


def _check_merged(oracle, naf, text, tmp=None, **kw):
    """the merged file must be indistinguishable, view by view, from a one-piece encode of the same text"""
    one = oracle.encode(text, **kw)[0]
    want = oracle.decode(one)                         # == text, except that FASTQ output never applies the mask
    assert oracle.decode(naf) == want
    for view in ("ids", "names", "lengths", "mask", "fasta", "total-length", "number"):
        assert oracle.decode(naf, view) == oracle.decode(one, view), view
    if kw.get("seq_type", "dna") in ("dna", "rna"):
        assert oracle.decode(naf, "4bit") == oracle.decode(one, "4bit")
    if helpers.have_ref():
        rc, out, err = helpers.ref_run("unnaf", [], naf)
        assert rc == 0 and out == want, err


CASES = [
    ("fastq", lambda: synth.fastq(301, 151, seed=3, lowercase=True, iupac=True), {}),
    ("fastq_odd", lambda: synth.fastq(77, 33, seed=4), {}),
    ("fasta_masked", lambda: synth.fasta_softmasked(200_001, width=60, seed=5, n_records=9, repeats=True, n_gaps=2), {}),
    ("ont", lambda: synth.ont_fasta(11, 1000, 9000, seed=6), {}),
    ("protein", lambda: synth.protein_fasta(300, 300, seed=7), {"seq_type": "protein"}),
    ("nomask", lambda: synth.fasta_softmasked(50_000, width=60, seed=8, n_records=5), {"no_mask": True}),
    ("all_lower", lambda: synth.fasta_reads(40, 150, seed=9).lower().replace(b">read", b">READ"), {}),
    ("empty_records", lambda: b">a\n>b\nACGTacgt\n>c\n>d x y\nacgtN\n>e\n", {}),
    ("fewer_records_than_ranks_fq", lambda: synth.fastq(3, 31, seed=10, lowercase=True), {}),
    ("one_record_fa", lambda: b">only\n" + b"acgtN" * 1001 + b"\n", {}),
]


@pytest.mark.parametrize("name,make,kw", CASES, ids=[c[0] for c in CASES])
def test_shards_local_oracle(oracle, name, make, kw):
    text = make()
    for world in (1, 2, 3, 5, 8):
        pieces = sharded.split_records(text, world)
        assert b"".join(pieces) == text and len(pieces) == world
        encs = [OracleShardEncoder(oracle, **kw) for _ in pieces]
        seq_type = helpers.SEQ_TYPES[kw.get("seq_type", "dna")]
        naf = sharded.encode_shards_local(encs, pieces, _opts(**kw), seq_type=seq_type)
        _check_merged(oracle, naf, text, **kw)


def test_link_math():
    C_ = sharded.Counts
    cs = [C_(n_records=1, n_bases=5, first_case=0, last_case=1, n_flips=1, last_flip=3, first_code=8),
          C_(),                                                              # empty shard in the middle
          C_(n_records=1, n_bases=4, first_case=1, last_case=1, first_code=2),
          C_(n_records=1, n_bases=3, first_case=0, last_case=0, first_code=4)]
    l0, l1, l2, l3 = (sharded.link_for(cs, r) for r in range(4))
    assert (l0.bases_before, l0.run_carry, l0.prev_last_case, l0.next_first_code, l0.is_last, l0.store_qual) == (0, 0, 0, 2, 0, 0)
    assert (l1.bases_before, l1.run_carry, l1.prev_last_case, l1.next_first_code) == (5, 2, 1, 2)
    assert (l2.bases_before, l2.run_carry, l2.prev_last_case, l2.next_first_code) == (5, 2, 1, 4)
    assert (l3.bases_before, l3.run_carry, l3.prev_last_case, l3.next_first_code, l3.is_last) == (9, 6, 1, 0, 1)


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
import helpers, test_sharded
from naf_b200 import api, sharded, synth
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
oracle = helpers.load_oracle()
ok = True
for name, make, kw in test_sharded.CASES[:4]:
    text = make()
    mine = sharded.split_records(text, 2)[dist.get_rank()]
    enc = test_sharded.OracleShardEncoder(oracle, **kw)
    out = sharded.encode_sharded(enc, mine, api.make_enc_opts(**kw), seq_type=helpers.SEQ_TYPES[kw.get("seq_type", "dna")])
    if dist.get_rank() == 0:
        test_sharded._check_merged(oracle, out.numpy().tobytes(), text, **kw)
    else:
        assert out is None
dist.barrier()
dist.destroy_process_group()
print("OK", flush=True)
"""


def test_shards_gloo_world2(tmp_path):
    """the real collective path (all_gather of counts and sizes, send / recv of the blocks) on CPU tensors"""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=helpers.ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0 and b"OK" in o, e.decode()[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name,make,kw", CASES, ids=[c[0] for c in CASES])
def test_shards_local_gpu(oracle, name, make, kw):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import naf_b200
    text = make()
    seq_type = helpers.SEQ_TYPES[kw.get("seq_type", "dna")]
    for world in (1, 2, 3, 5, 8):
        pieces = sharded.split_records(text, world)
        ctxs = [naf_b200.NafGpu(0) for _ in pieces]
        try:
            encs = [sharded.GpuShardEncoder(c) for c in ctxs]
            naf = sharded.encode_shards_local(encs, pieces, _opts(**kw), seq_type=seq_type)
            _check_merged(oracle, naf, text, **kw)
            assert ctxs[0].decode(naf) == oracle.decode(naf)
        finally:
            for c in ctxs:
                c.close()


@pytest.mark.gpu
def test_shards_gpu_bigger(oracle):
    """config-2 / config-5 shapes in 4 shards: multi-block streams per shard, odd nibble boundaries, long mask runs"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import naf_b200
    for text in (synth.fastq(60_001, 150, seed=51), synth.fasta_softmasked(6_000_001, width=60, seed=52, n_records=7, repeats=True, n_gaps=3)):
        pieces = sharded.split_records(text, 4)
        ctxs = [naf_b200.NafGpu(0) for _ in pieces]
        try:
            naf = sharded.encode_shards_local([sharded.GpuShardEncoder(c) for c in ctxs], pieces, _opts())
            assert ctxs[0].decode(naf) == text
            assert oracle.decode(naf) == text
            if helpers.have_ref():
                rc, out, err = helpers.ref_run("unnaf", [], naf)
                assert rc == 0 and out == text, err
        finally:
            for c in ctxs:
                c.close()


@pytest.mark.gpu
def test_record_range_decode(gpu, oracle):
    """one rank's piece of a multi-GPU decode: the pieces of any split concatenate to the full output, for files made by
    us (sequence-free frames: blocks outside the range are skipped) and by the reference (dependent blocks: decoded whole)"""
    files = []
    for text, kw in [(synth.fastq(40_001, 150, seed=61), {}), (synth.fasta_softmasked(4_000_001, width=60, seed=62, n_records=11, repeats=True), {}),
                     (synth.protein_fasta(30_000, 300, seed=63), {"seq_type": "protein"}), (synth.ont_fasta(60, 1000, 50000, seed=64), {})]:
        files.append(gpu.encode(text, **kw))
        if helpers.have_ref():
            args = ["--protein"] if kw.get("seq_type") == "protein" else []
            rc, naf, err = helpers.ref_run("ennaf", args + ["-c"], text)
            assert rc == 0, err
            files.append(naf)
    for case in helpers.manifest("cases"):
        files.append(helpers.golden("cases", case["name"] + ".naf"))
    for naf in files:
        for view in ("default", "fasta", "ids", "names", "sequences"):
            try:
                full = gpu.decode(naf, view)
            except Exception:
                continue
            for world in (1, 2, 3, 8):
                got = b"".join(sharded.decode_shard(gpu, naf, r, world, view) for r in range(world))
                assert got == full, (view, world, len(got), len(full))


def test_split_records_is_exact_on_ambiguous_fastq(oracle):
    """quality lines that begin with '@' (and '+' lines that repeat the name) must not fool the record splitter: it counts
    lines from the top instead of sniffing characters"""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(997):
        L = int(rng.integers(20, 90))
        seq = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)])
        qual = b"@" + bytes(rng.choice(np.frombuffer(b"@+IJ#", dtype=np.uint8), L - 1))
        recs.append(b"@r%d x\n" % i + seq + b"\n+r%d x\n" % i + qual + b"\n")
    text = b"".join(recs)
    for world in (2, 3, 7, 16):
        pieces = sharded.split_records(text, world)
        assert b"".join(pieces) == text and len(pieces) == world
        starts = set()
        off = 0
        for r in recs:
            starts.add(off); off += len(r)
        starts.add(len(text))
        off = 0
        for p in pieces:
            assert off in starts, "a piece does not start at a record boundary"
            off += len(p)
        naf = sharded.encode_shards_local([OracleShardEncoder(oracle) for _ in pieces], pieces, _opts())
        _check_merged(oracle, naf, text)


# ---- one device, file in pieces (sharded.encode_stream / iter_record_pieces / decode_stream)

def _chunks(text, size):
    return (text[i:i + size] for i in range(0, len(text), size))


@pytest.mark.parametrize("name,make,kw", CASES, ids=[c[0] for c in CASES])
def test_stream_pieces_and_encode_oracle(oracle, name, make, kw):
    """the piece cutter is exact (pieces concatenate to the input, every piece starts at a record), and the sequential
    two-encoder protocol gives a file indistinguishable from a one-piece encode"""
    text = make()
    seq_type = helpers.SEQ_TYPES[kw.get("seq_type", "dna")]
    for piece_bytes, chunk in ((700, 333), (5000, 4096), (len(text) // 3 + 1, 1 << 16), (1 << 30, 1000)):
        pieces = list(sharded.iter_record_pieces(_chunks(text, chunk), piece_bytes))
        assert b"".join(pieces) == text
        total = 0
        for p in pieces:
            assert p[:1] in (b">", b"@") or not p.strip()
            total += oracle.split(p, **kw)[1]["n_sequences"]
        assert total == oracle.split(text, **kw)[1]["n_sequences"]
        encs = [OracleShardEncoder(oracle, **kw) for _ in range(2)]
        naf = sharded.encode_stream(encs, iter(pieces), _opts(**kw), seq_type=seq_type)
        _check_merged(oracle, naf, text, **kw)


def test_stream_piece_without_bases(oracle):
    """a piece that holds only names (no base to complete the previous piece's last nibble) is merged into the next one"""
    text = b">a\nACG\n" + b">n1\n>n2\n>n3\n" * 40 + b">b\nTTGCA\n>c\n" + b">z\n" * 30
    for piece_bytes in (8, 20, 64):
        pieces = list(sharded.iter_record_pieces(_chunks(text, 7), piece_bytes))
        assert b"".join(pieces) == text and len(pieces) > 3
        naf = sharded.encode_stream([OracleShardEncoder(oracle) for _ in range(2)], iter(pieces), _opts())
        _check_merged(oracle, naf, text)
    assert oracle.decode(sharded.encode_stream([OracleShardEncoder(oracle) for _ in range(2)], iter([]), _opts())) == b""


@pytest.mark.gpu
def test_stream_gpu(oracle):
    """config-2 / config-5 shapes through two contexts on one GPU, ~1 MB pieces; decode_stream returns the text in ranges"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import naf_b200
    ctxs = [naf_b200.NafGpu(0) for _ in range(2)]
    try:
        for text in (synth.fastq(40_001, 150, seed=61, lowercase=True), synth.fasta_softmasked(5_000_001, width=60, seed=62, n_records=5, repeats=True, n_gaps=2)):
            pieces = list(sharded.iter_record_pieces(_chunks(text, 1 << 18), 1 << 20))
            assert b"".join(pieces) == text and len(pieces) >= 4
            naf = sharded.encode_stream([sharded.GpuShardEncoder(c) for c in ctxs], iter(pieces), _opts())
            want = oracle.decode(naf)
            assert want == oracle.decode(oracle.encode(text)[0])
            assert ctxs[0].decode(naf) == want
            if helpers.have_ref():
                rc, out, err = helpers.ref_run("unnaf", [], naf)
                assert rc == 0 and out == want, err
            got = []
            assert sharded.decode_stream(ctxs[1], naf, got.append, 5) == len(want)
            assert b"".join(got) == want
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.gpu
def test_record_cuts_on_the_gpu_equal_the_host_splitter(gpu):
    """nafgpu_record_cuts (newline ordinals by a prefix sum over tiles, one CTA per cut) against sharded.split_records: FASTQ
    whose quality lines begin with '@', FASTA with records far longer than a tile, leading white space, more pieces than records"""
    rng = np.random.default_rng(6)
    recs = []
    for i in range(2500):
        L = int(rng.integers(20, 300))
        seq = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)])
        qual = b"@" + bytes(rng.choice(np.frombuffer(b"@+IJ#>", dtype=np.uint8), L - 1))
        recs.append(b"@r%d x\n" % i + seq + b"\n+\n" + qual + b"\n")
    texts = [b"".join(recs), synth.fastq(5000, 150, seed=3), synth.fasta_softmasked(3_000_000, 60, seed=4, n_records=5),
             synth.ont_fasta(40, 10000, 50000, seed=5), b"\n \n" + synth.fasta_reads(50, 100, seed=6), b">only\nACGT\n", b""]
    for text in texts:
        for pieces in (1, 2, 3, 8, 64):
            want = [0]
            for p in sharded.split_records(text, pieces):
                want.append(want[-1] + len(p))
            want += [len(text)] * (pieces + 1 - len(want))
            assert sharded.split_records_gpu(gpu, text, pieces) == want, (len(text), pieces)
