"""Wall clock of bin/ennaf and bin/unnaf on one FASTQ file in /dev/shm: python tools/time_cli.py RECORDS [reps]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from naf_b200 import synth
import torch
torch.zeros(1).cuda()            # this process keeps a context alive, as a bench or a pipeline's other stage would: without it (and
                                 # without persistence mode) every tool start pays 1-2 s of driver initialisation on an idle GPU
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
work = f"/dev/shm/nafcli_{os.getpid()}"; os.makedirs(work, exist_ok=True)
fin, fnaf, fout = (os.path.join(work, x) for x in ("in.fq", "x.naf", "out.fq"))
text = synth.fastq(n, 150, seed=42)
open(fin, "wb").write(text)
for rep in range(reps):
    t0 = time.perf_counter(); subprocess.run([os.path.join(ROOT, os.environ.get("NAF_BIN", "bin"), "ennaf"), fin, "-o", fnaf], check=True)
    t1 = time.perf_counter(); subprocess.run([os.path.join(ROOT, os.environ.get("NAF_BIN", "bin"), "unnaf"), fnaf, "-o", fout], check=True)
    t2 = time.perf_counter()
    print(f"ennaf {t1 - t0:.3f} s  unnaf {t2 - t1:.3f} s", flush=True)
print("identical", open(fout, "rb").read() == text)
for f in os.listdir(work): os.remove(os.path.join(work, f))
os.rmdir(work)
