#!/bin/bash
# multi-GPU: parity tests + scaling bench (run with gpurun --gpus N)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_dist.py tests/test_gpu_spec.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_dist.txt; tail -5 gpurun_out/pytest_dist.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 448 --warmup 5 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
tail -c 2500 gpurun_out/bench_g$N.json; tail -5 gpurun_out/bench_g$N.err
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_g1.json'))
print('G1 value', b['value'], 'ms/step', b['ms_per_step'], 'e2e', b['e2e']['value'])
r=b['roofline']; print({k:round(r[k],5) for k in r if 'ms' in k})
PY
