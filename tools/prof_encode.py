"""Per-kernel times of encoding one BASELINE config: python tools/prof_encode.py [c2|c3|c4|c5] [scale]"""
import os, sys
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
for rep in range(3):
    ctx.profile(rep == 2)
    ctx.encode_device(d.data_ptr(), t.size, api.make_enc_opts(**kw))
print(cfg, "text", t.size, "kernels_ms", round(ctx.timing().kernels_ms, 2), "launches", ctx.timing().kernel_launches, "fallback", ctx.timing().parser_fallback)
for name, cnt, ms in sorted(ctx.profile_report(), key=lambda x: -x[2])[:14]:
    print(f"  {name:22s} x{cnt:4d} {ms:8.3f} ms")
