# tuning helper: forest kernel variants on the bench workload (run under gpurun)
for cfg in ${CFGS:-pipe32 pipe16 dyn32 dyn16}; do
NLF_FOREST_CFG=$cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_$cfg.json 2> gpurun_out/b_$cfg.err
python -c "
import json; d=json.load(open('gpurun_out/b_$cfg.json')); print('$cfg', round(d['value']/1e6,1), 'Mpx/s', round(d['ms_per_step'],2), 'forest', round(d['kernels']['forest']['ms_per_step'],2), d['kernels']['forest']['launches_per_step'], 'feat', round(d['kernels']['feature']['ms_per_step'],2))" || tail -2 gpurun_out/b_$cfg.err
done
