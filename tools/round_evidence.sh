#!/bin/bash
# The measurements a round is judged on, in one GPU call: tools/round_evidence.sh TAG   (outputs: gpurun_out/TAG_*)
T=${1:-rX}
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -2 gpurun_out/${T}_bench_reference.err
# launch list of the bench command (per-launch times are cold-cache and serialised: the SHARES are what counts)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_2Mreads.csv \
    python bench.py --records 2000000 --steps 2 --warmup 3 --no-cpu-baseline --c5-gbp 0 > gpurun_out/${T}_launches_bench.log 2>&1
wc -l gpurun_out/${T}_launches_2Mreads.csv
