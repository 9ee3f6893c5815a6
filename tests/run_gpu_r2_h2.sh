#!/bin/bash
# Round 2, 2-GPU call on the final code (gpurun --gpus 2, charged 2x): sharded parity on real NVLink, the driver's bench
# command at N=2 (fused exchange) with its parity leg, and a C5 rehearsal (2 x 2,097,152 nodes, device-side churn)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider -k "p2p" 2>&1 | tail -4 | tee gpurun_out/r2h${N}_pytest_dist.txt
show() { python - "$1" <<'PY'
import json, sys
try:
    b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = b['roofline']; t = r.get('timeline') or {}
    print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s parity %s exch %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches'), b.get('parity_check'), b['config'].get('exchange')))
    print('    windows', b['timing']['windows_ms'])
    print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run_bench() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
      bench.py --gpus $N --no-cpu "$@" > gpurun_out/r2h${N}_$tag.json 2> gpurun_out/r2h${N}_$tag.err
  tail -2 gpurun_out/r2h${N}_$tag.err | cut -c1-300
  show gpurun_out/r2h${N}_$tag.json
}
run_bench bench20 --steps 20 --warmup 5
run_bench bench448
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29729 \
    studies/c5_suspicion_sweep.py --nodes-per-gpu 2097152 --rounds 300 --suspicion 3 8 --sample-every 50 \
    > gpurun_out/r2h${N}_c5.jsonl 2> gpurun_out/r2h${N}_c5.err
tail -3 gpurun_out/r2h${N}_c5.err | cut -c1-300
cut -c1-900 gpurun_out/r2h${N}_c5.jsonl
