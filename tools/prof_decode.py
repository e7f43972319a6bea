import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers, naf_b200
from naf_b200 import synth
n = int(sys.argv[1]); kind = sys.argv[2]; iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
naf_path = f"/dev/shm/prof_{kind}_{n}.naf"
if not os.path.exists(naf_path):
    text = synth.fastq(n, 150, seed=1) if kind == "fastq" else synth.fasta_softmasked(n, 60, 1, 3, True, 4)
    rc, naf, err = helpers.ref_run("ennaf", ["-c"], text, tmp="/dev/shm")
    open(naf_path, "wb").write(naf)
naf = open(naf_path, "rb").read()
ctx = naf_b200.NafGpu(0)
for i in range(iters):
    ctx.decode_raw(naf, naf_b200.api.make_dec_opts())
    tm = ctx.timing(); print("kernels ms", tm.kernels_ms, "launches", tm.kernel_launches)
