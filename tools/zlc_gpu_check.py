"""The data-parallel LZ stage (NAFGPU_LZ=shared: k_zlc_find / k_zlc_define / k_zlc_finish) on a B200, in one short process:
  A. frames of single streams equal, byte for byte, what the CPU emulation of the same HD bodies writes (tests/emu/emu_zlzc.cpp),
     and decode back on the device
  B. a whole FASTQ / FASTA file: the oracle and the device decode the .naf back to the text; sizes next to level 1
  C. per-kernel times of one encode of N reads (CUDA events), and the encode / decode call times with and without the stage
python tools/zlc_gpu_check.py [reads_for_timing] [log] [NAFGPU_LZ for parts A and B: shared | b]      (no torch; every line is flushed, so a cut-off run still reports)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["NAFGPU_LZ"] = sys.argv[3] if len(sys.argv) > 3 else "shared"      # "b": the finder's bit-mask formulation (same frames)
T0 = time.time()
LOG = open(sys.argv[2], "a") if len(sys.argv) > 2 else None


def say(**kw):
    kw["t"] = round(time.time() - T0, 2)
    line = json.dumps(kw)
    print(line, flush=True)
    if LOG:
        LOG.write(line + "\n"); LOG.flush()


def main():
    import struct
    import numpy as np
    import naf_b200
    from naf_b200 import synth
    import helpers
    n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    import shutil
    emu = os.path.join(ROOT, "tests", "_build", "emu_zlzc_bytes")
    if shutil.which("g++"):                                   # from the sources of this snapshot; else the binary that travelled
        emu = os.path.join(os.environ.get("TMPDIR", "/tmp"), "emu_zlzc_%d" % os.getpid())
        subprocess.run(["g++", "-std=c++17", "-O2", "-DZLC_BYTES", "-o", emu, os.path.join(ROOT, "tests", "emu", "emu_zlzc.cpp")], check=True)
    ctx = naf_b200.NafGpu(0)
    say(step="context")
    oracle = helpers.load_oracle()
    rng = np.random.default_rng(9)
    text = synth.fastq(100_000, 150, seed=5)
    streams, _ = oracle.split(text)
    ont = synth.ont_fasta(300, 10000, 30000, seed=6)
    ostreams, _ = oracle.split(ont)
    cases = {"ids": streams[0], "comments": streams[1], "lengths": streams[2], "ont_ids": ostreams[0], "ont_comments": ostreams[1],
             "ont_lengths": ostreams[2], "ont_mask": ostreams[3],
             "rle_between": b"x" * 9000 + b"".join(b"SRR1.%d\0" % i for i in range(5, 3005)) + b"y" * 20000 + b"".join(b"SRR1.%d\0" % i for i in range(77777, 80777)),
             "raw_first": bytes(rng.integers(0, 256, 9000, dtype=np.uint8)) + b"".join(b"SRR1.%d\0" % i for i in range(123, 4123)),
             "unseen_bytes": b"".join(b"SRR1.%d\0" % i for i in range(1, 3001)) + bytes(rng.integers(0, 256, 20000, dtype=np.uint8)) + b"".join(b"SRR1.%d\0" % i for i in range(9, 3009)),
             "noise4": bytes(rng.integers(0, 4, 70000, dtype=np.uint8)), "empty": b"", "one": b"a", "rle": b"q" * 30000,
             "runs": b"".join(bytes([65 + (i * i) % 23]) * (1 + i % 40) for i in range(3000)),
             "units150": struct.pack("<I", 150) * 50000, "short": b"SRR1.1\0SRR1.2\0SRR1.3\0"}
    tmp = os.environ.get("TMPDIR", "/tmp")
    ok_all = True
    for name, data in cases.items():
        inp, z = os.path.join(tmp, "zlc_i.bin"), os.path.join(tmp, "zlc_o.zst")
        with open(inp, "wb") as f:
            f.write(data)
        subprocess.run([emu, inp, z, "8192"], check=True, capture_output=True)
        want = open(z, "rb").read()
        got = ctx.zstd_compress(data, level=2)
        same = got == want
        back = ctx.zstd_decompress(got, len(data)) == data
        ok_all &= same and back
        say(step="A", case=name, raw=len(data), frame=len(got), equals_emulation=same, device_decodes=back)
    for label, t, kw in (("fastq", text, {}), ("ont_fasta", ont, {})):
        naf = ctx.encode(t, **kw)
        o_ok = oracle.decode(naf) == t
        d_ok = ctx.decode(naf) == t
        ok_all &= o_ok and d_ok
        say(step="B", case=label, text=len(t), naf=len(naf), ratio=round(len(naf) / len(t), 4), oracle_decodes=o_ok, device_decodes=d_ok)
    say(step="verdict", all_ok=bool(ok_all))
    if not ok_all:
        sys.exit(1)
    if n_time <= 0:
        return
    big = synth.fastq(n_time, 150, seed=42)
    say(step="C", made_reads=n_time, text=len(big))
    for mode in ("0", "shared", "b"):                          # "b": the finder as bit masks;                               # "0": the level-1 parse (entropy only), same process, same box
        os.environ["NAFGPU_LZ"] = mode
        naf = ctx.encode(big)
        ctx.profile(True)
        naf = ctx.encode(big)
        erep = sorted(ctx.profile_report(), key=lambda x: -x[2])
        out = ctx.decode(naf)
        drep = sorted(ctx.profile_report(), key=lambda x: -x[2])
        ctx.profile(False)
        say(step="C", mode=mode, naf=len(naf), ratio=round(len(naf) / len(big), 4), roundtrip_ok=out == big,
            encode_kernels={n: round(ms, 3) for n, c, ms in erep[:9]}, decode_kernels={n: round(ms, 3) for n, c, ms in drep[:9]})
        enc, dec = [], []
        for _ in range(4):
            naf = ctx.encode(big)
            enc.append(round(ctx.timing().kernels_ms, 2))
            out = ctx.decode(naf)
            dec.append(round(ctx.timing().kernels_ms, 2))
        say(step="C", mode=mode, encode_kernels_ms=enc, decode_kernels_ms=dec)


if __name__ == "__main__":
    try:
        main()
    except Exception as e:                                    # noqa: BLE001 -- report and fail
        say(step="error", error=repr(e)[:400])
        raise
