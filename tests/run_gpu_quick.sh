#!/bin/bash
# quick single-GPU check: parity tests + bench (no ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
python bench.py --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err
SWIM_SPLIT=1 python bench.py --no-cpu > gpurun_out/bench_q_nograph.json 2>> gpurun_out/bench_q.err
python bench.py --no-cpu --warmup 600 --steps 448 > gpurun_out/bench_q_quiet.json 2>> gpurun_out/bench_q.err
python - <<'PY'
import json
for f in ['bench_q','bench_q_nograph','bench_q_quiet']:
    try:
        b=json.load(open(f'gpurun_out/{f}.json'))
        r=b['roofline']
        print(f, 'value %.3e'%b['value'], 'us/round %.2f'%(b['ms_per_step']*1e3), 'e2e %.3e'%b['e2e']['value'], 'launches', b['gpu_launches'], {k:round(r[k]*1e3,2) for k in r if 'ms' in k}, b['convergence'])
    except Exception as e: print(f, 'ERR', e)
PY
