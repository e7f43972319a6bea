"""Host-buffer (e2e) encode + decode through the C ABI, pinned buffers, wall clock per call: python tools/ab_e2e.py RECORDS [reps]
(run once per build: NAFGPU_LIB=path python tools/ab_e2e.py ...)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, naf_b200
from naf_b200 import api, synth

n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
a = synth.fastq_array(n, 150, seed=42)
h_text = torch.from_numpy(a).pin_memory()
n_text = h_text.numel()
ctx = naf_b200.NafGpu(0)
eo, do = api.make_enc_opts(), api.make_dec_opts()
enc, dec = [], []
h_naf = torch.empty(n_text // 2 + 4096, dtype=torch.uint8).pin_memory()
for it in range(reps + 2):
    t0 = time.perf_counter()
    addr, size, info = ctx.encode_raw((h_text.data_ptr(), n_text), eo)
    t1 = time.perf_counter()
    C.memmove(h_naf.data_ptr(), addr, size)
    t2 = time.perf_counter()
    taddr, tsize = ctx.decode_raw((h_naf.data_ptr(), size), do)
    t3 = time.perf_counter()
    if it == 0:
        got = (C.c_uint8 * 64).from_address(taddr + tsize - 64)
        assert tsize == n_text and bytes(got) == bytes(a[-64:].tobytes())
    if it >= 2:
        enc.append((t1 - t0) * 1e3); dec.append((t3 - t2) * 1e3)
print(json.dumps({"lib": os.environ.get("NAFGPU_LIB", "default"), "encode_ms": [round(x, 1) for x in enc], "decode_ms": [round(x, 1) for x in dec],
                  "e2e_gbases_s": round(n * 150 / ((min(enc) + min(dec)) * 1e-3) / 1e9, 2)}))
