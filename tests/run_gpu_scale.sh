#!/bin/bash
# scaling check on an N-GPU box: dist parity tests, then bench at 1, 2, .. N GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_dist.txt; tail -3 gpurun_out/pytest_dist.txt
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu > gpurun_out/scale_g1.json 2> gpurun_out/scale_g1.err
for G in 2 4 8; do
  if [ $G -le $N ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2971$G bench.py --gpus $G > gpurun_out/scale_g$G.json 2> gpurun_out/scale_g$G.err
  fi
done
python - <<'PY'
import json, glob
base=None
for g in (1,2,4,8):
    try:
        b=json.loads(open(f'gpurun_out/scale_g{g}.json').read().strip().splitlines()[-1])
    except Exception as e:
        continue
    if g==1: base=b['value']
    r=b['roofline']
    print(f"G{g} value {b['value']:.3e} us/round {b['ms_per_step']*1e3:.2f} eff {b['value']/(g*base) if base else 0:.2f} e2e {b['e2e']['value']:.3e} conv {b['convergence']['rounds_to_convergence']} exch {b['config'].get('exchange')}", {k:round(r[k]*1e3,2) for k in r if 'ms' in k})
PY
