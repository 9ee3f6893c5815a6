#!/bin/bash
# Round 2, call C (1 GPU): sender-side membership filter (K2 sees only envelopes that can matter), C3 full-size parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2c_pytest_gpu.txt; tail -4 gpurun_out/r2c_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = b['roofline']; t = r.get('timeline') or {}
print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches')))
print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'))
print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
print('    split', {k: round(v, 2) for k, v in r['split_kernels_us'].items()}, 'frac %.3f' % r['frac'], 'conv', (b.get('convergence') or {}).get('rounds_to_convergence'))
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench20.json 2> gpurun_out/r2c_bench20.err; tail -3 gpurun_out/r2c_bench20.err; show gpurun_out/r2c_bench20.json
timeout 600 python bench.py --no-cpu > gpurun_out/r2c_bench448.json 2> gpurun_out/r2c_bench448.err; show gpurun_out/r2c_bench448.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches20.csv python bench.py --steps 20 --warmup 5 --no-cpu --windows 1 --spinup 0 > gpurun_out/r2c_ncu.log 2>&1
grep -c . gpurun_out/r2c_launches20.csv
