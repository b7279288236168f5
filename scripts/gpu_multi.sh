set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_peer_gather.py 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; tail -c 300 gpurun_out/scale_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    tail -c 300 gpurun_out/scale_n$n.json; tail -5 gpurun_out/scale_n$n.err
    IDKPT_GATHER=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_nccl_n$n.json 2> gpurun_out/scale_nccl_n$n.err
  fi
done
