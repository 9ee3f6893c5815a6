#!/bin/bash
# 8-GPU box: world-8 parity tests, then bench at 4 and 8 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -m gpu -q -k "small and 8" 2>&1 | tail -30 > gpurun_out/pytest_dist8.txt; tail -3 gpurun_out/pytest_dist8.txt
for G in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2972$G bench.py --gpus $G --converge-limit 700 > gpurun_out/scale_g$G.json 2> gpurun_out/scale_g$G.err
  tail -c 600 gpurun_out/scale_g$G.err | tail -4
done
python - <<'PY'
import json
base=5.09e10
for g in (4,8):
    try:
        b=json.loads(open(f'gpurun_out/scale_g{g}.json').read().strip().splitlines()[-1])
        r=b['roofline']
        print(f"G{g} value {b['value']:.3e} us/round {b['ms_per_step']*1e3:.2f} eff {b['value']/(g*base):.2f} e2e {b['e2e']['value']:.3e} conv {b['convergence']} exch {b['config'].get('exchange')}", {k:round(r[k]*1e3,2) for k in r if 'ms' in k})
    except Exception as e: print(g, 'ERR', e)
PY
