#!/bin/bash
# One gpurun session; steps selected by arguments (default: tests probe). Everything lands in gpurun_out/.
#   tests   pytest -m gpu            probe   scripts/variant_probe.py --run --eighth      sweep   probe + threshold sweep
#   bench   python bench.py          ncu     launch list + --set full capture of the bench
#   slots   strict multi-GPU parity in one process (scripts/check_global_slots.py)   slotsN  bench.py --global-slots on all GPUs
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
    config4N) N=$(nvidia-smi -L | wc -l); timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload config4 --steps 10 > gpurun_out/config4_n$N.json 2> gpurun_out/config4_n$N.err; tail -3 gpurun_out/config4_n$N.err; cut -c1-600 gpurun_out/config4_n$N.json;;
    config4) timeout 900 python bench.py --workload config4 --steps 10 --no-cpu-baseline --no-vxgi > gpurun_out/config4_n1.json 2> gpurun_out/config4_n1.err; tail -3 gpurun_out/config4_n1.err; cut -c1-400 gpurun_out/config4_n1.json;;
    config3) timeout 900 python bench.py --workload config3 --steps 10 --no-cpu-baseline --no-vxgi > gpurun_out/config3_n1.json 2> gpurun_out/config3_n1.err; tail -3 gpurun_out/config3_n1.err; cut -c1-400 gpurun_out/config3_n1.json;;
    vxgiN) N=$(nvidia-smi -L | wc -l); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/check_vxgi_multigpu.py > gpurun_out/vxgi_n$N.json 2> gpurun_out/vxgi_n$N.err; tail -3 gpurun_out/vxgi_n$N.err; cat gpurun_out/vxgi_n$N.json;;
    ref) timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json;;
    ncu) # launch list of the bench command (shares of the step) + --set full captures of the dominant kernels (never a bench number)
       timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-vxgi > gpurun_out/bench_under_ncu.log 2>&1
       timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traverse2 -s 8 -c 3 -f -o gpurun_out/prof_traverse2_r02 python scripts/ncu_target.py pt > gpurun_out/ncu_t2.log 2>&1
       timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 2 -f -o gpurun_out/prof_shade_r02 python scripts/ncu_target.py pt > gpurun_out/ncu_shade.log 2>&1
       timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compact -s 8 -c 2 -f -o gpurun_out/prof_compact_r02 python scripts/ncu_target.py pt > gpurun_out/ncu_compact.log 2>&1
       timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vx_ -c 14 -f -o gpurun_out/prof_vxgi_r02 python scripts/ncu_target.py vxgi > gpurun_out/ncu_vx.log 2>&1
       ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_t2.log;;
    sanitize) timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log; tail -6 gpurun_out/sanitize_memcheck.log
       timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.log; tail -4 gpurun_out/sanitize_racecheck.log;;
    slots) for W in 2 3; do CUDA_DEVICE_MAX_CONNECTIONS=32 IDKPT_GATHER_TIMEOUT_MS=2000 timeout 200 python scripts/check_global_slots.py --world $W > gpurun_out/global_slots_w$W.json 2> gpurun_out/global_slots_w$W.err; rc=$?; echo "slots w$W exit $rc"; tail -3 gpurun_out/global_slots_w$W.err; cat gpurun_out/global_slots_w$W.json
         if [ $rc -ne 0 ]; then CUDA_MODULE_LOADING=EAGER CUDA_DEVICE_MAX_CONNECTIONS=32 IDKPT_GATHER_TIMEOUT_MS=2000 timeout 200 python scripts/check_global_slots.py --world $W > gpurun_out/global_slots_w${W}_eager.json 2> gpurun_out/global_slots_w${W}_eager.err; echo "slots w$W eager exit $?"; tail -3 gpurun_out/global_slots_w${W}_eager.err; cat gpurun_out/global_slots_w${W}_eager.json; break; fi; done;;
    slotsN) N=$(nvidia-smi -L | wc -l); timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --global-slots > gpurun_out/bench_slots_n$N.json 2> gpurun_out/bench_slots_n$N.err; tail -5 gpurun_out/bench_slots_n$N.err; cut -c1-1500 gpurun_out/bench_slots_n$N.json;;
    smoke) timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3;;
    *) echo "unknown step $s";;
  esac
done
