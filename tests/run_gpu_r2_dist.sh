#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus 2, charged 2x): the sharded launch paths against each other
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tests/run_gpu_r2_dist.sh 2'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run_bench() { # $1 = tag, env comes from the caller
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
      bench.py --gpus $N --no-cpu > gpurun_out/r2_dist_$1.json 2> gpurun_out/r2_dist_$1.err
  python - "$1" <<'PY'
import json, sys
try:
    b = json.loads(open('gpurun_out/r2_dist_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    r = b['roofline']
    print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value']),
          {k: round(r[k] * 1e3, 2) for k in r if 'ms' in k})
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  tail -2 gpurun_out/r2_dist_$1.err
}
# default: split launches + peer_barrier_kernel (round 1: 41 us/round at 2 GPUs)
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
run_bench split
# opt-in: round_kernel + grid_peer_barrier, multi-round launches
SWIM_ROUND_KERNEL=1 timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider -k p2p 2>&1 | tail -4
SWIM_ROUND_KERNEL=1 run_bench roundkernel
ls -la gpurun_out | head
