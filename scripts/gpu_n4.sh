set -x
mkdir -p gpurun_out
for n in 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/n${n}_detail.json 2> gpurun_out/n${n}_detail.err
tail -3 gpurun_out/n${n}_detail.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n${n}_detail.json').read().strip().splitlines()[-1])
print('N=$n', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['kernel_ms_per_step'])
for r in d['per_rank_ms']: print(r)
PY
done
