#!/bin/bash
# multi-GPU: parity tests + scaling bench (run with gpurun --gpus N)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_dist.py tests/test_gpu_spec.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_dist.txt; tail -5 gpurun_out/pytest_dist.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 448 --warmup 5 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
tail -c 2500 gpurun_out/bench_g$N.json; tail -5 gpurun_out/bench_g$N.err
