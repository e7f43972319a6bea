#!/bin/bash
# ncu --set full capture of one kernel of the encode path: tools/prof_encode_ncu.sh <kernel regex> <out name> [cfg] [scale]
K=${1:-k_fused}; O=${2:-prof}; CFG=${3:-c2}; S=${4:-0.1}
ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/$O python tools/prof_encode.py $CFG $S > gpurun_out/$O.log 2>&1
tail -20 gpurun_out/$O.log
