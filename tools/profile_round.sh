#!/bin/bash
# One gpurun call that refreshes everything under profiles/ for a round:
#   bash tools/profile_round.sh r01      (run on the GPU box; results land in gpurun_out/)
# Each ncu pass runs only after the plain command exited 0.
R=${1:-r01}
O=gpurun_out/$R
mkdir -p $O
set -x
python bench.py > $O/bench.json 2> $O/bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain.json 2> $O/plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_forest_pipe|k_features|k_diag_means|k_isotonic|k_threshold|k_pool|k_count_candidates" -s 16 -c 10 \
    -o $O/full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O
