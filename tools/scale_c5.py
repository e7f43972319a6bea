#!/usr/bin/env python3
"""BASELINE configs[4]: ONE soft-masked human-like FASTA (24 records, line width 60) decoded by N ranks (strong scaling).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_c5.py [--gbp 3.0]

Rank 0 makes the text and encodes it on its GPU (untimed); the .naf (a quarter of the text) is broadcast over NCCL and
every rank keeps it in HBM with a host mirror, as it would after reading the file.  Timed: every rank decodes its share of
the records (nafgpu_decode_device with first_record / n_records), device-resident, max over ranks.  Checked: the pieces'
sizes and byte sums against rank 0's text.  One JSON line from rank 0."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gbp", type=float, default=3.0)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import naf_b200
    from naf_b200 import api, sharded, synth, container

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = naf_b200.NafGpu(local)
    cudart = C.CDLL("libcudart.so")
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    n_bases = int(a.gbp * 1e9)
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    d_text = None
    if rank == 0:
        text = np.frombuffer(synth.fasta_softmasked(n_bases, 60, seed=42, n_records=24, repeats=True, n_gaps=20), dtype=np.uint8)
        d_text = torch.zeros(text.size + 64, dtype=torch.uint8, device=dev)
        d_text[:text.size] = torch.from_numpy(text.copy()).to(dev)
        addr, size, info = ctx.encode_device(d_text.data_ptr(), text.size, api.make_enc_opts())
        d_naf = torch.empty(size, dtype=torch.uint8, device=dev)
        cudart.cudaMemcpy(d_naf.data_ptr(), addr, size, 3)
        meta[0], meta[1] = size, text.size
    dist.broadcast(meta, src=0)
    if rank != 0:
        d_naf = torch.empty(int(meta[0]), dtype=torch.uint8, device=dev)
    dist.broadcast(d_naf, src=0)
    h_naf = d_naf.cpu()
    n_rec = 24
    first, count = sharded.record_range(n_rec, rank, world)
    opts = api.make_dec_opts(first_record=first, n_records=count)
    times, tsize, taddr = [], 0, 0
    for rep in range(a.reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if count:
            taddr, tsize = ctx.decode_device(d_naf.data_ptr(), d_naf.numel(), (h_naf.data_ptr(), h_naf.numel()), opts)
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
    piece = torch.empty(tsize, dtype=torch.uint8, device=dev)
    if tsize:
        cudart.cudaMemcpy(piece.data_ptr(), taddr, tsize, 3)
    mine = torch.tensor([tsize, int(piece.sum(dtype=torch.int64).item()) if tsize else 0], dtype=torch.int64, device=dev)
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    tmin = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
    dist.all_reduce(tmin, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok, off = True, 0
        for r in range(world):
            sz, sm = int(allp[r][0]), int(allp[r][1])
            ok = ok and int(d_text[off:off + sz].sum(dtype=torch.int64).item()) == sm
            off += sz
        ok = ok and off == int(meta[1])
        print(json.dumps({"config": "c5", "workload": f"{a.gbp} Gbp soft-masked FASTA, 24 records, one .naf", "n_gpus": world, "naf_bytes": int(meta[0]),
                          "text_bytes": int(meta[1]), "decode_ms": round(float(tmin.item()) * 1e3, 3), "decode_gbases_s": round(n_bases / float(tmin.item()) / 1e9, 2),
                          "pieces_verified": ok, "scaling": "strong"}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
