#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N, charged Nx): sharded parity + bench on the fused path (default) and the split path
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tests/run_gpu_r2_f2.sh 2'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
show() { python - "$1" <<'PY'
import json, sys
try:
    b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = b['roofline']; t = r.get('timeline') or {}
    print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s parity %s exch %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches'), b.get('parity_check'), b['config'].get('exchange')))
    print('    windows', b['timing']['windows_ms'])
    print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
    print('    split', {k: round(v, 2) for k, v in r['split_kernels_us'].items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run_bench() { # $1 = tag, $2.. = bench args; env from the caller
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
      bench.py --gpus $N --no-cpu "$@" > gpurun_out/r2f2_${N}_$tag.json 2> gpurun_out/r2f2_${N}_$tag.err
  tail -2 gpurun_out/r2f2_${N}_$tag.err | cut -c1-300
  show gpurun_out/r2f2_${N}_$tag.json
}
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider -k "p2p" 2>&1 | tail -4
run_bench rk20 --steps 20 --warmup 5
run_bench rk448
SWIM_ROUND_KERNEL=0 run_bench split20 --steps 20 --warmup 5
