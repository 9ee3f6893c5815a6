#!/bin/bash
# Round 2, last call (gpurun --gpus 2): the driver's bench command at N=2. It stopped behind the parity leg — the end-to-end
# probe raced an unsynchronised swim_sim_load against a peer's first publication; bench.py fixed afterwards
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k2_bench20.json 2> gpurun_out/r2k2_bench20.err
grep "^\[bench" gpurun_out/r2k2_bench20.err | cut -c1-220
python - <<'PY'
import json
try:
    b = json.loads(open('gpurun_out/r2k2_bench20.json').read().strip().splitlines()[-1])
    print('2 GPUs: value %.3e us/round %.2f e2e %.3e parity %s exch %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('parity_check'), b['config'].get('exchange')), b['timing']['windows_ms'], b['e2e'].get('windows_ms'))
    print('   conv', b.get('convergence'))
except Exception as e:
    print('bench FAILED', e)
PY
tail -3 gpurun_out/r2k2_bench20.err | cut -c1-300
