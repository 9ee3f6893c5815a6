#!/bin/bash
# Round 2, 8-GPU call (gpurun --gpus 8, charged 8x): world-8 parity, the weak-scaling point, and BASELINE config C5 itself:
# N = 16,777,216 nodes on 8 GPUs, 10 % churn (crash 1e-3 per round, rejoin U[10,50]), suspicion timeout sweep S = 2,3,5,8,13
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
# (world-8 parity: bench.py's parity_check leg below compares all 8 shards with the oracle at 8 x C3)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29728 \
    bench.py --gpus 8 --no-cpu --steps 20 --warmup 5 --converge-limit 400 > gpurun_out/r2g8_bench20.json 2> gpurun_out/r2g8_bench20.err
tail -2 gpurun_out/r2g8_bench20.err | cut -c1-300
python - <<'PY'
import json
try:
    b = json.loads(open('gpurun_out/r2g8_bench20.json').read().strip().splitlines()[-1])
    print('8 GPUs: value %.3e us/round %.2f e2e %.3e parity %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('parity_check')), b['timing']['windows_ms'])
except Exception as e:
    print('bench FAILED', e)
PY
ROUNDS=${1:-1000}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29729 \
    studies/c5_suspicion_sweep.py --nodes-per-gpu 2097152 --rounds $ROUNDS --suspicion 2 3 5 8 13 --sample-every 50 \
    > gpurun_out/r2g8_c5.jsonl 2> gpurun_out/r2g8_c5.err
tail -3 gpurun_out/r2g8_c5.err | cut -c1-300
cut -c1-700 gpurun_out/r2g8_c5.jsonl
