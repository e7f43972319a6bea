"""Per-kernel times of decoding a REFERENCE-made .naf (oracle/_ref/ennaf on the host, then nafgpu_decode_device):
python tools/prof_refdecode.py [c2|c3|c4|c5] [scale]"""
import ctypes as C, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, naf_b200
from naf_b200 import api, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"; s = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
text = {"c2": lambda: synth.fastq(int(10_000_000 * s), 150, seed=42), "c3": lambda: synth.ont_fasta(int(100_000 * s), 10000, 50000, seed=42),
        "c4": lambda: synth.protein_fasta(int(1_000_000 * s), 300, seed=42), "c5": lambda: synth.fasta_softmasked(int(3_000_000_000 * s), 60, seed=42, n_records=24, repeats=True, n_gaps=20)}[cfg]()
work = f"/dev/shm/refdec_{os.getpid()}"; os.makedirs(work, exist_ok=True)
open(work + "/in.txt", "wb").write(text)
subprocess.run([ROOT + "/oracle/_ref/ennaf"] + (["--protein"] if cfg == "c4" else []) + [work + "/in.txt", "-o", work + "/ref.naf"], check=True, env=dict(os.environ, TMPDIR=work))
naf = np.fromfile(work + "/ref.naf", dtype=np.uint8)
for f in os.listdir(work): os.remove(work + "/" + f)
os.rmdir(work)
d = torch.zeros(naf.size + 64, dtype=torch.uint8, device="cuda"); d[:naf.size] = torch.from_numpy(naf).cuda(); h = torch.from_numpy(naf)
ctx = naf_b200.NafGpu(0)
for rep in range(3):
    ctx.profile(rep == 2)
    t = time.perf_counter(); ctx.decode_device(d.data_ptr(), naf.size, (h.data_ptr(), naf.size), api.make_dec_opts()); wall = time.perf_counter() - t
print(cfg, "text", len(text), "naf", naf.size, "wall ms", round(wall * 1e3, 2), "kernels_ms", round(ctx.timing().kernels_ms, 2), "launches", ctx.timing().kernel_launches)
for name, cnt, ms in sorted(ctx.profile_report(), key=lambda x: -x[2])[:14]:
    print(f"  {name:22s} x{cnt:4d} {ms:8.3f} ms")
