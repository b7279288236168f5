set -x
mkdir -p gpurun_out
export IDKPT_TUNE_SETUP=8 IDKPT_TUNE_LEAF=4
ncu --set full --clock-control none --import-source on -k regex:k_traverse2 -s 45 -c 3 -o gpurun_out/prof_traverse2_r1_b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1; tail -2 gpurun_out/ncu_full_b.log
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 45 -c 3 -o gpurun_out/prof_shade_r1_b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c.log 2>&1; tail -2 gpurun_out/ncu_full_c.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --height 136 > gpurun_out/bench_eighth.json 2>gpurun_out/bench_eighth.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_eighth.json').read().strip().splitlines()[-1])
print('EIGHTH', round(d['value'],1), d['ms_per_step'], d['kernel_ms_per_step'], d['wall_ms_per_step'], d['gpu_launches'])
"
