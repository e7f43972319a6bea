#!/usr/bin/env python3
"""ennaf / unnaf across the GPUs of one box: ONE .naf file, N ranks (naf_b200/sharded.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/naf_multi.py encode IN.fq -o OUT.naf [--protein ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/naf_multi.py decode IN.naf -o OUT.fq

encode: rank 0 finds record-aligned cut points (exact: '>' at a line start for FASTA, newline counting for 4-line FASTQ)
and broadcasts them; every rank encodes its piece; one all-gather of counts, one gather of zstd blocks over NCCL; rank 0
writes the file.  The result is a plain .naf: the reference unnaf (and bin/unnaf) read it.
decode: every rank reads the (small) .naf, decodes its share of the records and writes its bytes at their offset in the
output file."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["encode", "decode"])
    ap.add_argument("input")
    ap.add_argument("-o", "--output", required=True)
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

    import numpy as np
    import torch
    import torch.distributed as dist
    import naf_b200
    from naf_b200 import api, sharded

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = naf_b200.NafGpu(local)
    data = np.fromfile(a.input, dtype=np.uint8)

    if a.mode == "encode":
        cuts = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        if rank == 0:                                       # record boundaries from the GPU scan (nafgpu_record_cuts)
            cuts[:] = torch.tensor(sharded.split_records_gpu(ctx, data, world), dtype=torch.int64)
        dist.broadcast(cuts, src=0)
        lo, hi = int(cuts[rank]), int(cuts[rank + 1])
        mine = data[lo:hi].tobytes()
        opts = api.make_enc_opts(seq_type=a.seq_type, no_mask=a.no_mask, line_length=a.line_length)
        out = sharded.encode_sharded(sharded.GpuShardEncoder(ctx), mine, opts, seq_type=api._SEQ_TYPES[a.seq_type],
                                     title=a.title.encode() if a.title else None, line_length=a.line_length, device=dev)
        if rank == 0:
            out.cpu().numpy().tofile(a.output)
    else:
        text = sharded.decode_shard(ctx, data.tobytes(), rank, world, a.view, no_mask=a.no_mask, line_length=a.line_length)
        size = torch.tensor([len(text)], dtype=torch.int64, device=dev)
        sizes = [torch.empty_like(size) for _ in range(world)]
        dist.all_gather(sizes, size)
        offset = sum(int(s) for s in sizes[:rank])
        if rank == 0:
            open(a.output, "wb").close()
        dist.barrier()
        fd = os.open(a.output, os.O_WRONLY)
        os.pwrite(fd, text, offset)
        os.close(fd)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
