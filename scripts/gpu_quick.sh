set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('QUICK', round(d['value'],1), 'Mrays/s e2e', round(d['e2e']['value'],1), d['kernel_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], 'wall', d['wall_ms_per_step'], 'cpu', d.get('cpu_baseline',{}).get('value'))
"
