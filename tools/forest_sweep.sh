# tuning helper: forest kernel variants on the bench workload (run under gpurun)
timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -2
for cfg in ${CFGS:-flat4x8 flat4x4 flat4x6 flat4x12 flat3x8 flat5x8 coop16x1}; do
NLF_FOREST_CFG=$cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_$cfg.json 2> gpurun_out/b_$cfg.err
python -c "
import json; d=json.load(open('gpurun_out/b_$cfg.json')); print('$cfg', round(d['value']/1e6,1), 'Mpx/s', round(d['ms_per_step'],2), 'forest', round(d['kernels']['forest']['ms_per_step'],2), d['kernels']['forest']['launches_per_step'], 'feat', round(d['kernels']['feature']['ms_per_step'],2))" || tail -2 gpurun_out/b_$cfg.err
done
