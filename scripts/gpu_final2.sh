set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('FINAL', round(d['value'],1), 'serial', round(d['rays_per_step']/d['serial_ms_per_step']/1e3,1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
timeout 300 python scripts/textured_bench.py 2>&1 | tail -4
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -c 200 gpurun_out/bench_under_ncu.log
