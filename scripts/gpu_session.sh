#!/bin/bash
# One gpurun session; steps selected by arguments (default: tests probe). Everything lands in gpurun_out/.
#   tests   pytest -m gpu            probe   scripts/variant_probe.py --run --eighth      sweep   probe + threshold sweep
#   bench   python bench.py          ncu     launch list + --set full capture of the bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/session_gpu.txt 2>&1
steps="${@:-tests probe}"
for s in $steps; do
  case $s in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -16 gpurun_out/pytest_gpu.log;;
    probe) timeout 900 python scripts/variant_probe.py --run --eighth > gpurun_out/variant_probe.log 2>&1; cut -c1-900 gpurun_out/variant_probe.log;;
    sweep) timeout 900 python scripts/variant_probe.py --run --eighth --sweep > gpurun_out/variant_probe.log 2>&1; cut -c1-900 gpurun_out/variant_probe.log;;
    bench) timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json;;
    benchN) N=$(nvidia-smi -L | wc -l); timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json;;
    ref) timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json;;
    *) echo "unknown step $s";;
  esac
done
