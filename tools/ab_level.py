"""A/B of the encoder's parse levels on one GPU, device-resident (CUDA events on the library's stream):
    python tools/ab_level.py RECORDS [OUT.json]
level 1 = entropy-only parse of every stream; level 2 = LZ77 + FSE-coded sequences on ids / comments / lengths, in each of its
formulations (NAFGPU_LZ is read per call): "2" the data-parallel stage a level selects, "2b" the same with the finder's phases as
bit masks, "2t" the thread-per-block stage of round 1 (also covers the mask).
Each: 2 warm-up + 3 timed round trips (bit-exact check on the device), then one profiled step (per-kernel ms)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, naf_b200
from naf_b200 import api, synth

n = int(sys.argv[1]); out_path = sys.argv[2] if len(sys.argv) > 2 else None
text = torch.from_numpy(synth.fastq_array(n, 150, seed=42))
n_text = text.numel()
d_text = torch.zeros(n_text + 64, dtype=torch.uint8, device="cuda"); d_text[:n_text] = text.cuda()
ctx = naf_b200.NafGpu(0)
stream = torch.cuda.ExternalStream(ctx.lib.nafgpu_stream(ctx.h))
cudart = C.CDLL("libcudart.so"); cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
res = {"records": n, "text_bytes": n_text, "levels": {}}
d_naf = torch.zeros(n_text // 2 + 4096, dtype=torch.uint8, device="cuda")
for level, env in ((1, None), ("2", None), ("2b", "b"), ("2t", "1")):
    if env is None:
        os.environ.pop("NAFGPU_LZ", None)
    else:
        os.environ["NAFGPU_LZ"] = env
    eo, do = api.make_enc_opts(level=int(str(level)[0])), api.make_dec_opts()
    enc, dec = [], []
    for it in range(5):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        addr, size, info = ctx.encode_device(d_text.data_ptr(), n_text, eo)
        ev[1].record(stream)
        torch.cuda.synchronize()
        cudart.cudaMemcpy(d_naf.data_ptr(), addr, size, 3)
        h_naf = d_naf[:size].cpu()
        torch.cuda.synchronize()
        ev[2].record(stream)
        taddr, tsize = ctx.decode_device(d_naf.data_ptr(), size, (h_naf.data_ptr(), size), do)
        ev[3].record(stream)
        torch.cuda.synchronize()
        if it == 0:
            back = torch.empty(tsize, dtype=torch.uint8, device="cuda")
            cudart.cudaMemcpy(back.data_ptr(), taddr, tsize, 3)
            assert tsize == n_text and torch.equal(back, d_text[:n_text]), f"level {level}: round trip differs"
            del back
        if it >= 2:
            enc.append(ev[0].elapsed_time(ev[1])); dec.append(ev[2].elapsed_time(ev[3]))
    ctx.profile(True)
    ctx.encode_device(d_text.data_ptr(), n_text, eo)
    prof = {nm: ms for nm, c, ms in ctx.profile_report()}
    ctx.decode_device(d_naf.data_ptr(), size, (h_naf.data_ptr(), size), do)
    for nm, c, ms in ctx.profile_report():
        prof[nm] = prof.get(nm, 0.0) + ms
    ctx.profile(False)
    e, d = sum(enc) / len(enc), sum(dec) / len(dec)
    res["levels"][str(level)] = {
        "naf_bytes": int(size), "ratio": size / n_text, "stream_comp": [int(x) for x in info.stream_comp],
        "encode_ms": e, "decode_ms": d, "encode_gbases_s": n * 150 / e / 1e6, "decode_gbases_s": n * 150 / d / 1e6,
        "roundtrip_gbases_s": n * 150 / (e + d) / 1e6,
        "kernels_ms": {k: round(v, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:24]},
    }
    print(json.dumps({str(level): res["levels"][str(level)]}), flush=True)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)
