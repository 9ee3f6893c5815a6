#!/bin/bash
# Round 2, call A (1 GPU): where HEAD stands on hardware — GPU suite, smoke, the driver's bench command, the long window, quiet-batch A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader | head -2
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2a_pytest_gpu.txt; tail -6 gpurun_out/r2a_pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r2a_smoke.txt
show() { python - "$1" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = b['roofline']
print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches')))
print('   ', {k: round(r[k] * 1e3, 2) for k in r if k.endswith('ms_per_launch')}, 'conv', (b.get('convergence') or {}).get('rounds_to_convergence'))
PY
}
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2a_bench20_$i.json 2> gpurun_out/r2a_bench20_$i.err; show gpurun_out/r2a_bench20_$i.json
done
timeout 600 python bench.py --no-cpu > gpurun_out/r2a_bench448.json 2> gpurun_out/r2a_bench448.err; show gpurun_out/r2a_bench448.json
for Q in 0 8; do
  SWIM_QUIET_BATCH=$Q timeout 300 python bench.py --no-cpu > gpurun_out/r2a_bench448_qb$Q.json 2> gpurun_out/r2a_bench448_qb$Q.err; show gpurun_out/r2a_bench448_qb$Q.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches20.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2a_ncu.log 2>&1
grep -c . gpurun_out/r2a_launches20.csv
ls -la gpurun_out | head -30
