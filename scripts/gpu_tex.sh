set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_textures.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/time_post.py 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tex.json 2> gpurun_out/bench_tex.err; tail -3 gpurun_out/bench_tex.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_tex.json').read().strip().splitlines()[-1])
print('QUICK', round(d['value'],1), 'Mrays/s e2e', round(d['e2e']['value'],1), d['kernel_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
"
