set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 24 --warmup 3 > gpurun_out/n4_async.json 2> gpurun_out/n4_async.err
tail -3 gpurun_out/n4_async.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n4_async.json').read().strip().splitlines()[-1])
print('N=4', round(d['value'],1), 'serial', round(d['rays_per_step']/d['serial_ms_per_step']/1e3,1), 'e2e', round(d['e2e']['value'],1), [round(r['pipelined_total'],3) for r in d['per_rank_ms']], d['clocks'])
PY
