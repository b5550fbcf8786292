#!/bin/bash
# One gpurun call that refreshes everything under profiles/ for a round:
#   bash tools/profile_round.sh r02      (run on the GPU box; results land in gpurun_out/<round>/)
# Each ncu pass runs only after the plain command exited 0.  The ncu passes use a quarter of the default workload
# (48 of the 192 cfg3 assemblies, same shapes): ncu replays every kernel ~40 times.
R=${1:-r02}
O=gpurun_out/$R
mkdir -p $O
set -x
python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py > $O/bench.json 2> $O/bench.err || exit 1
SMALL="--steps 2 --warmup 3 --no-cpu-baseline --no-cli --assemblies 48"
python bench.py $SMALL > $O/plain.json 2> $O/plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py $SMALL > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_forest_pipe|k_features|k_diag_means|k_isotonic|k_threshold|k_pool|k_count_candidates" -s 18 -c 12 \
    -o $O/full python bench.py $SMALL > $O/ncu_full.log 2>&1
NLF_PIPELINE_DEPTH=8 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cli > $O/depth8.json 2> $O/depth8.err
ls -la $O
