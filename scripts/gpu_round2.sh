set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; tail -c 2500 gpurun_out/bench_r1_c.json; tail -3 gpurun_out/bench_r1_c.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_c_reference.json 2>&1; tail -c 700 gpurun_out/bench_r1_c_reference.json
timeout 600 python scripts/run_configs.py config5 sort > gpurun_out/configs_a.jsonl 2> gpurun_out/configs_a.err; cat gpurun_out/configs_a.jsonl | cut -c1-1500; tail -3 gpurun_out/configs_a.err
timeout 900 python scripts/run_configs.py config3 config4 > gpurun_out/configs_b.jsonl 2> gpurun_out/configs_b.err; cat gpurun_out/configs_b.jsonl | cut -c1-1500; tail -3 gpurun_out/configs_b.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_c.log 2>&1; tail -2 gpurun_out/ncu_launch_c.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 27 -c 2 -o gpurun_out/prof_shade_r1_c python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_shade.log 2>&1; tail -4 gpurun_out/ncu_full_shade.log
ls -la gpurun_out | tail -12
