set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_peer_gather.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi or peer or gather" 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 24 --warmup 3 > gpurun_out/n2_async.json 2> gpurun_out/n2_async.err
tail -3 gpurun_out/n2_async.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n2_async.json').read().strip().splitlines()[-1])
print('N=2', round(d['value'],1), 'serial', round(d['rays_per_step']/d['serial_ms_per_step']/1e3,1), 'e2e', round(d['e2e']['value'],1), d['per_rank_ms'])
PY
timeout 300 python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/n1_async.json 2> gpurun_out/n1_async.err; tail -2 gpurun_out/n1_async.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n1_async.json').read().strip().splitlines()[-1])
print('N=1', round(d['value'],1), 'serial', round(d['rays_per_step']/d['serial_ms_per_step']/1e3,1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
PY
