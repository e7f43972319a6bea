#!/usr/bin/env python3
"""ennaf / unnaf on ONE GPU for files that do not fit in HBM: the file goes through the device in record-aligned pieces
(naf_b200/sharded.py: encode_stream / decode_stream).

    python tools/naf_stream.py encode IN.fq -o OUT.naf [--piece-mb 512] [--protein ...]     (IN may be '-': stdin)
    python tools/naf_stream.py decode IN.naf -o OUT.fq [--pieces 8]                          (OUT may be '-': stdout)

encode: two contexts on the GPU take the pieces alternately (piece i is parsed while piece i-1, which needs piece i's
first base for its last nibble, is finished and its zstd blocks are fetched); the blocks of each stream accumulate on
the host and the container is written at the end -- the reference does the same with one temp file per stream
(ennaf.c:538-589).  The result is a plain .naf.
decode: the records are decoded in `--pieces` consecutive ranges, each written as soon as it is done."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_chunks(f, size=1 << 24):
    while True:
        b = f.read(size)
        if not b:
            return
        yield b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["encode", "decode"])
    ap.add_argument("input")
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--piece-mb", type=int, default=512)
    ap.add_argument("--pieces", type=int, default=8)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--dna", dest="seq_type", action="store_const", const="dna", default="dna")
    ap.add_argument("--rna", dest="seq_type", action="store_const", const="rna")
    ap.add_argument("--protein", dest="seq_type", action="store_const", const="protein")
    ap.add_argument("--text", dest="seq_type", action="store_const", const="text")
    ap.add_argument("--no-mask", action="store_true")
    ap.add_argument("--title")
    ap.add_argument("--line-length", type=int)
    ap.add_argument("--fasta", dest="view", action="store_const", const="fasta", default="default")
    ap.add_argument("--fastq", dest="view", action="store_const", const="fastq")
    a = ap.parse_args()

    import naf_b200
    from naf_b200 import api, sharded

    fin = sys.stdin.buffer if a.input == "-" else open(a.input, "rb")
    fout = sys.stdout.buffer if a.output == "-" else open(a.output, "wb")
    if a.mode == "encode":
        ctxs = [naf_b200.NafGpu(a.device) for _ in range(2)]
        opts = api.make_enc_opts(seq_type=a.seq_type, no_mask=a.no_mask, line_length=a.line_length, level=a.level)
        pieces = sharded.iter_record_pieces(read_chunks(fin), a.piece_mb << 20)
        naf = sharded.encode_stream([sharded.GpuShardEncoder(c) for c in ctxs], pieces, opts, seq_type=api._SEQ_TYPES[a.seq_type],
                                    title=a.title.encode() if a.title else None, line_length=a.line_length)
        fout.write(naf)
    else:
        ctx = naf_b200.NafGpu(a.device)
        sharded.decode_stream(ctx, fin.read(), fout.write, a.pieces, a.view, no_mask=a.no_mask, line_length=a.line_length)
    fout.flush()


if __name__ == "__main__":
    main()
