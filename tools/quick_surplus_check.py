"""GPU decode of files with surplus sequence (ennaf's id-byte bug) against the oracle; prints one line per case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers, naf_b200
o = helpers.load_oracle()
gpu = naf_b200.NafGpu(0)
cases = [(b">>\na", "text"), (b">a\x01b\nACGTAC\nGG\n>c\x02\x03\nTTTTTTT\n", "dna"),
         (b">a\x01b x\nacgtnnac\nGG\n>c\x02\x03\nTTTTTTT\n>e\n>f\n", "dna"), (b">p\x01\nMKV\nLLA\n>q\x7f\x01\x01\nMM\n", "protein"),
         (b"@r\x01\nACGT\n+\nIIII\n", "dna"), (b">u\x01\x01\x01\nACGU\n" + b">v\x02\nacguACGU\n" * 300, "rna")]
bad = 0
for text, st in cases:
    for maker, naf in (("oracle", o.encode(text, seq_type=st)[0]), ("gpu", gpu.encode(text, seq_type=st)), ("gpu-l3", gpu.encode(text, seq_type=st, level=3))):
        for kw in ({}, {"line_length": 3}, {"line_length": 0}, {"no_mask": True}):
            want = o.decode(naf, "fasta", **kw)
            try:
                got = gpu.decode(naf, "fasta", **kw)
            except Exception as e:
                got = repr(e).encode()
            ok = got == want
            bad += not ok
            print("OK  " if ok else "DIFF", st, maker, kw, "" if ok else (got[-40:], want[-40:]), flush=True)
print("bad =", bad, flush=True)
