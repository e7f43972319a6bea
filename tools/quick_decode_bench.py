import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers, naf_b200
from naf_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "fastq"
t = time.time()
text = synth.fastq(n, 150, seed=1) if kind == "fastq" else synth.fasta_softmasked(n, 60, 1, 3, True, 4)
bases = n * 150 if kind == "fastq" else n
print("gen", time.time() - t, len(text))
t = time.time(); rc, naf, err = helpers.ref_run("ennaf", ["-c"], text, tmp="/dev/shm"); tref_e = time.time() - t
print("ref ennaf", tref_e, len(naf), "Gbases/s", bases / tref_e / 1e9)
open("/dev/shm/x.naf", "wb").write(naf)
t = time.time(); rc, out, err = helpers.ref_run("unnaf", ["/dev/shm/x.naf", "-o", "/dev/shm/x.out"]); tref_d = time.time() - t
print("ref unnaf", tref_d, "Gbases/s", bases / tref_d / 1e9)
ctx = naf_b200.NafGpu(0)
for i in range(4):
    t = time.time(); out = ctx.decode(naf); dt = time.time() - t
    tm = ctx.timing()
    print(f"gpu decode wall {dt:.4f}s  h2d {tm.h2d_ms:.2f} kernels {tm.kernels_ms:.2f} d2h {tm.d2h_ms:.2f} total {tm.total_ms:.2f} ms launches {tm.kernel_launches}  -> {bases/tm.kernels_ms/1e6:.1f} Gbases/s kernels, {bases/tm.total_ms/1e6:.1f} e2e")
assert out == text
print("OK identical")
