set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "1 8 8" "2 8 8" "2 4 8" "2 12 8" "2 8 4" "2 8 12" "2 8 16" "2 4 4" "2 16 16" "2 1 1"; do
  set -- $cfg
  IDKPT_TRAVERSE_VARIANT=$1 IDKPT_TUNE_SETUP=$2 IDKPT_TUNE_LEAF=$3 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/tune.json 2> gpurun_out/tune.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/tune.json').read().strip().splitlines()[-1])
print('TUNE variant=$1 setup=$2 leaf=$3', round(d['value'],1), 'Mrays/s', d['kernel_ms_per_step'], 'frac', round(d['roofline']['frac'],3))
"
done
