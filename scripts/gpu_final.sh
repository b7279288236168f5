set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; tail -3 gpurun_out/bench_r1_f.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r1_f.json').read().strip().splitlines()[-1])
print('FINAL', round(d['value'],1), 'Mrays/s e2e', round(d['e2e']['value'],1), d['kernel_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'cpu', d.get('cpu_baseline',{}).get('value'))
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_f.log 2>&1; tail -2 gpurun_out/ncu_launch_f.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_traverse2 -s 40 -c 2 -o gpurun_out/prof_traverse2_r1_f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_f1.log 2>&1; tail -2 gpurun_out/ncu_full_f1.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_compact" -s 90 -c 4 -o gpurun_out/prof_shade_r1_f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_f2.log 2>&1; tail -2 gpurun_out/ncu_full_f2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vx -c 14 -o gpurun_out/prof_vxgi_r1_f python scripts/run_configs.py config5 > gpurun_out/ncu_full_f3.log 2>&1; tail -2 gpurun_out/ncu_full_f3.log
ls -la gpurun_out | tail -8
