"""A small piped encode + decode of every input shape at both levels, for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import naf_b200
from naf_b200 import synth
ctx = naf_b200.NafGpu(0)
os.environ.update({"NAFGPU_PIPE_MIN": str(1 << 20), "NAFGPU_PIPE_CHUNK": str(1 << 18), "NAFGPU_PIPE_PIECE": str(1 << 18)})
for name, text, kw in [("fastq", synth.fastq(20000, 150, seed=31), {}), ("fasta", synth.fasta_softmasked(3_000_000, 60, seed=33, n_records=5, repeats=True, n_gaps=2), {}),
                       ("crlf", synth.fasta_reads(3000, 150, seed=1).replace(b"\n", b"\r\n"), {}), ("protein", synth.protein_fasta(4000, 300, seed=35), {"seq_type": "protein"})]:
    naf = ctx.encode(text, **kw)
    out = ctx.decode(naf)
    assert out == text.replace(b"\r\n", b"\n"), name
    naf2 = ctx.encode(text, level=2, **kw)
    assert ctx.decode(naf2) == out, name
    print(name, len(text), len(naf), len(naf2), flush=True)
print("ok")
