"""Per-kernel times of decoding our own .naf of one BASELINE config: python tools/prof_decode_cfg.py [c2|c3|c4|c5] [scale]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, naf_b200
from naf_b200 import api, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"; s = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
text, kw = {"c2": lambda: (synth.fastq(int(10_000_000 * s), 150, seed=42), {}), "c3": lambda: (synth.ont_fasta(int(100_000 * s), 10000, 50000, seed=42), {}),
            "c4": lambda: (synth.protein_fasta(int(1_000_000 * s), 300, seed=42), {"seq_type": "protein"}),
            "c5": lambda: (synth.fasta_softmasked(int(3_000_000_000 * s), 60, seed=42, n_records=24, repeats=True, n_gaps=20), {})}[cfg]()
t = np.frombuffer(text, dtype=np.uint8)
d = torch.zeros(t.size + 64, dtype=torch.uint8, device="cuda"); d[:t.size] = torch.from_numpy(t.copy()).cuda()
ctx = naf_b200.NafGpu(0)
cudart = C.CDLL("libcudart.so"); cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
addr, size, info = ctx.encode_device(d.data_ptr(), t.size, api.make_enc_opts(**kw))
dn = torch.zeros(size + 64, dtype=torch.uint8, device="cuda"); cudart.cudaMemcpy(dn.data_ptr(), addr, size, 3); hn = dn[:size].cpu()
for rep in range(3):
    ctx.profile(rep == 2)
    ctx.decode_device(dn.data_ptr(), size, (hn.data_ptr(), size), api.make_dec_opts())
print(cfg, "text", t.size, "naf", size, "kernels_ms", round(ctx.timing().kernels_ms, 2), "launches", ctx.timing().kernel_launches)
for name, cnt, ms in sorted(ctx.profile_report(), key=lambda x: -x[2])[:12]:
    print(f"  {name:22s} x{cnt:4d} {ms:8.3f} ms")
