#!/bin/bash
# Round 2, 8-GPU call (gpurun --gpus 8, charged 8x): BASELINE config C5 itself — N = 16,777,216 nodes on 8 GPUs, churn (crash
# 1e-3 per node and round, rejoin U[10,50]), suspicion-timeout sweep S = 2,3,5,8,13 on uniform-random views —, then the
# driver's bench command at N=8 with its parity leg (all 8 shards vs the oracle at 8 x C3), then C5 with ring-lattice views
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
ROUNDS=${1:-1000}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29729 \
    studies/c5_suspicion_sweep.py --nodes-per-gpu 2097152 --rounds $ROUNDS --suspicion 2 3 5 8 13 --sample-every 100 \
    > gpurun_out/r2h8_c5_random.jsonl 2> gpurun_out/r2h8_c5_random.err
tail -2 gpurun_out/r2h8_c5_random.err | cut -c1-300
grep "^{" gpurun_out/r2h8_c5_random.jsonl | cut -c1-500
timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29728 \
    bench.py --gpus 8 --no-cpu --steps 20 --warmup 5 --converge-limit 520 > gpurun_out/r2h8_bench20.json 2> gpurun_out/r2h8_bench20.err
grep "^\[bench" gpurun_out/r2h8_bench20.err | cut -c1-200 | tail -12
python - <<'PY'
import json
try:
    b = json.loads(open('gpurun_out/r2h8_bench20.json').read().strip().splitlines()[-1])
    print('8 GPUs: value %.3e us/round %.2f e2e %.3e parity %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('parity_check')), b['timing']['windows_ms'])
    t = b['roofline'].get('timeline') or {}
    print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
except Exception as e:
    print('bench FAILED', e)
PY
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29730 \
    studies/c5_suspicion_sweep.py --nodes-per-gpu 2097152 --rounds $ROUNDS --suspicion 3 8 --sample-every 100 --topology ring \
    > gpurun_out/r2h8_c5_ring.jsonl 2> gpurun_out/r2h8_c5_ring.err
tail -2 gpurun_out/r2h8_c5_ring.err | cut -c1-300
grep "^{" gpurun_out/r2h8_c5_ring.jsonl | cut -c1-500
