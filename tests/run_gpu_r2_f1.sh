#!/bin/bash
# Round 2, call F1 (1 GPU): receive folded into the next scan phase (2 barriers per busy round), probes_per_round, ring workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2f_pytest_gpu.txt; tail -5 gpurun_out/r2f_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = b['roofline']; t = r.get('timeline') or {}
print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s parity %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches'), b.get('parity_check')))
print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'))
print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
print('    split', {k: round(v, 2) for k, v in r['split_kernels_us'].items()}, 'frac %.3f' % r['frac'])
g = b.get('state_machine_workload')
if g: print('    ring: value %.3e us/round %.2f conv %s applied/sent %.3f parity %s' % (g['value'], g['ms_per_step']*1e3, g['rounds_to_convergence'], g['recs_applied_over_recs_sent'], g['parity_check']), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (g.get('timeline') or {}).items() if k != 'what'})
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench20.json 2> gpurun_out/r2f_bench20.err; tail -3 gpurun_out/r2f_bench20.err; show gpurun_out/r2f_bench20.json
timeout 600 python bench.py --no-cpu > gpurun_out/r2f_bench448.json 2> gpurun_out/r2f_bench448.err; tail -3 gpurun_out/r2f_bench448.err; show gpurun_out/r2f_bench448.json
