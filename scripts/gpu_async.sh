set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_async.py -x -q 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for L in 4 2 3 6; do
IDKPT_LANES=$L timeout 300 python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lanes$L.json 2> gpurun_out/bench_lanes$L.err; tail -2 gpurun_out/bench_lanes$L.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_lanes$L.json').read().strip().splitlines()[-1])
print('LANES=$L', round(d['value'],1), 'serial', round(d['rays_per_step']/d['serial_ms_per_step']/1e3,1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))
PY
done
