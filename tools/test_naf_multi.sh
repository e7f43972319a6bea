#!/bin/bash
# smoke test of tools/naf_multi.py on N GPUs (default 2): one file in, one .naf out, reference unnaf and our multi-rank decode agree
set -e
N=${1:-2}
W=/dev/shm/nafmulti_$$; mkdir -p $W
python - <<PY
import sys; sys.path.insert(0, ".")
from naf_b200 import synth
open("$W/a.fq", "wb").write(synth.fastq(300001, 150, seed=5))
open("$W/b.fa", "wb").write(synth.fasta_softmasked(30000001, 60, seed=6, n_records=7, repeats=True, n_gaps=2))
PY
for f in a.fq b.fa; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/naf_multi.py encode $W/$f -o $W/$f.naf 2>&1 | grep -v -i warn | tail -2
  TMPDIR=$W oracle/_ref/unnaf $W/$f.naf -o $W/$f.ref.out
  cmp $W/$f $W/$f.ref.out && echo "$f: reference unnaf reproduces the input from the $N-rank .naf"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/naf_multi.py decode $W/$f.naf -o $W/$f.multi.out 2>&1 | grep -v -i warn | tail -2
  cmp $W/$f $W/$f.multi.out && echo "$f: $N-rank decode reproduces the input"
  ls -l $W/$f $W/$f.naf | awk '{print $5, $9}'
done
rm -rf $W
