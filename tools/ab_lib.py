"""A/B of two builds of libnafgpu.so on the bench workload, device-resident: python tools/ab_lib.py RECORDS [level]
(run once per build: NAFGPU_LIB=path python tools/ab_lib.py ...).  Prints encode / decode ms (CUDA events, 3 timed after 2
warm-up round trips, bit-exact check on the device) and the per-kernel times of one profiled step."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, naf_b200
from naf_b200 import api, synth

n = int(sys.argv[1]); level = int(sys.argv[2]) if len(sys.argv) > 2 else 1
text = torch.from_numpy(synth.fastq_array(n, 150, seed=42))
n_text = text.numel()
d_text = torch.zeros(n_text + 64, dtype=torch.uint8, device="cuda"); d_text[:n_text] = text.cuda()
ctx = naf_b200.NafGpu(0)
stream = torch.cuda.ExternalStream(ctx.lib.nafgpu_stream(ctx.h))
cudart = C.CDLL("libcudart.so"); cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
d_naf = torch.zeros(n_text // 2 + 4096, dtype=torch.uint8, device="cuda")
eo, do = api.make_enc_opts(level=level), api.make_dec_opts()
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
        assert tsize == n_text and torch.equal(back, d_text[:n_text]), "round trip differs"
        del back
    if it >= 2:
        enc.append(ev[0].elapsed_time(ev[1])); dec.append(ev[2].elapsed_time(ev[3]))
ctx.profile(True)
ctx.encode_device(d_text.data_ptr(), n_text, eo)
prof = {nm: ms for nm, c, ms in ctx.profile_report()}
ctx.decode_device(d_naf.data_ptr(), size, (h_naf.data_ptr(), size), do)
for nm, c, ms in ctx.profile_report():
    prof[nm] = prof.get(nm, 0.0) + ms
e, d = sum(enc) / len(enc), sum(dec) / len(dec)
print(json.dumps({"lib": os.environ.get("NAFGPU_LIB", "default"), "naf_bytes": int(size), "encode_ms": round(e, 3), "decode_ms": round(d, 3),
                  "roundtrip_gbases_s": round(n * 150 / (e + d) / 1e6, 2),
                  "kernels_ms": {k: round(v, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:10]}}), flush=True)
