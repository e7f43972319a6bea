#!/usr/bin/env python3
"""Full-size round trips of the BASELINE.json configs on one B200: bit-exact check + device-resident throughput.

    python tools/config_sweep.py [--scale 1.0] [--configs c2,c3,c4,c5] > gpurun_out/configs.json

For each config: generate the synthetic text (naf_b200.synth, fixed seeds), encode on the device, decode on the
device, compare the decoded text with the input ON THE DEVICE (size-independent parity property at full size:
decode(encode(x)) == x; FASTQ inputs here are upper-case, so the mask quirk does not apply), and time both directions
with the library's CUDA events (kernels only, inputs resident in HBM).  One JSON line per config."""
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
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="c2,c3,c4,c5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ref", action="store_true", help="also: the UNMODIFIED reference ennaf (oracle/_ref) makes the .naf on the host, "
                    "the GPU decodes it; must be byte-identical to what the reference unnaf prints (= the input here)")
    a = ap.parse_args()
    import numpy as np
    import torch
    import naf_b200
    from naf_b200 import api, synth
    s = a.scale
    gens = {
        "c2": ("10M x 150 bp FASTQ", lambda: synth.fastq_array(int(10_000_000 * s), 150, seed=42), int(10_000_000 * s) * 150, {}),
        "c3": ("100k x 10-50 kbp ONT-like FASTA, IUPAC + soft-masked", lambda: np.frombuffer(synth.ont_fasta(int(100_000 * s), 10000, 50000, seed=42), dtype=np.uint8), None, {}),
        "c4": ("1M x 300 aa protein FASTA", lambda: np.frombuffer(synth.protein_fasta(int(1_000_000 * s), 300, seed=42), dtype=np.uint8), int(1_000_000 * s) * 300, {"seq_type": "protein"}),
        "c5": ("3 Gbp soft-masked FASTA, 24 records, width 60", lambda: np.frombuffer(synth.fasta_softmasked(int(3_000_000_000 * s), 60, seed=42, n_records=24, repeats=True, n_gaps=20), dtype=np.uint8), int(3_000_000_000 * s), {}),
    }
    ctx = naf_b200.NafGpu(0)
    cudart = C.CDLL("libcudart.so")
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    for name in a.configs.split(","):
        what, gen, bases, kw = gens[name]
        t0 = time.time()
        text = gen()
        n = int(text.size)
        d_text = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
        d_text[:n] = torch.from_numpy(np.ascontiguousarray(text)).cuda()
        del text
        gen_s = time.time() - t0
        eo, do = api.make_enc_opts(**kw), api.make_dec_opts()
        enc_ms, dec_ms, ok, naf_size, fallback = [], [], True, 0, 0
        for rep in range(a.reps):
            addr, size, info = ctx.encode_device(d_text.data_ptr(), n, eo)
            t = ctx.timing(); enc_ms.append(t.kernels_ms); fallback = t.parser_fallback
            if bases is None:
                bases = int(info.n_bases)
            d_naf = torch.zeros(size + 64, dtype=torch.uint8, device="cuda")
            cudart.cudaMemcpy(d_naf.data_ptr(), addr, size, 3)
            h_naf = d_naf[:size].cpu()
            taddr, tsize = ctx.decode_device(d_naf.data_ptr(), size, (h_naf.data_ptr(), size), do)
            dec_ms.append(ctx.timing().kernels_ms)
            out = torch.empty(tsize, dtype=torch.uint8, device="cuda")
            cudart.cudaMemcpy(out.data_ptr(), taddr, tsize, 3)
            ok = ok and tsize == n and bool(torch.equal(out, d_text[:n]))
            naf_size = size
            del out, d_naf
        e, d = min(enc_ms), min(dec_ms)
        ref = None
        ref_ennaf = os.path.join(ROOT, "oracle", "_ref", "ennaf")
        if a.ref and os.access(ref_ennaf, os.X_OK):
            import subprocess
            work = f"/dev/shm/nafsweep_{os.getpid()}"
            os.makedirs(work, exist_ok=True)
            d_text[:n].cpu().numpy().tofile(os.path.join(work, "in.txt"))
            t1 = time.time()
            args = [ref_ennaf] + (["--protein"] if kw.get("seq_type") == "protein" else []) + [os.path.join(work, "in.txt"), "-o", os.path.join(work, "ref.naf")]
            subprocess.run(args, check=True, env=dict(os.environ, TMPDIR=work))
            ref_enc_s = time.time() - t1
            rnaf = np.fromfile(os.path.join(work, "ref.naf"), dtype=np.uint8)
            d_rnaf = torch.zeros(rnaf.size + 64, dtype=torch.uint8, device="cuda")
            d_rnaf[:rnaf.size] = torch.from_numpy(rnaf).cuda()
            h_rnaf = torch.from_numpy(rnaf)
            rms, rok = [], True
            for rep in range(a.reps):
                taddr, tsize = ctx.decode_device(d_rnaf.data_ptr(), rnaf.size, (h_rnaf.data_ptr(), rnaf.size), do)
                rms.append(ctx.timing().kernels_ms)
                out = torch.empty(tsize, dtype=torch.uint8, device="cuda")
                cudart.cudaMemcpy(out.data_ptr(), taddr, tsize, 3)
                rok = rok and tsize == n and bool(torch.equal(out, d_text[:n]))
                del out
            ref = {"reference_naf_bytes": int(rnaf.size), "reference_ennaf_s": round(ref_enc_s, 1), "gpu_decode_of_reference_naf_bit_exact": rok,
                   "gpu_decode_of_reference_naf_ms": round(min(rms), 3), "gpu_decode_of_reference_naf_gbases_s": round(bases / min(rms) / 1e6, 2)}
            for f in os.listdir(work):
                os.remove(os.path.join(work, f))
            os.rmdir(work)
            del d_rnaf
        print(json.dumps({"config": name, "workload": what, "reference_made_file": ref, "text_bytes": n, "bases": bases, "naf_bytes": naf_size, "ratio": round(naf_size / n, 4),
                          "round_trip_bit_exact": ok, "fast_parser": not fallback, "encode_ms": round(e, 3), "decode_ms": round(d, 3),
                          "encode_gbases_s": round(bases / e / 1e6, 2), "decode_gbases_s": round(bases / d / 1e6, 2),
                          "host_generation_s": round(gen_s, 1)}), flush=True)
        del d_text
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
