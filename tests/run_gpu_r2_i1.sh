#!/bin/bash
# Round 2, call I1 (1 GPU): K1b draws in one Philox pass + lane-held proxies, K1a risky nodes compacted over the warp,
# churn -> events -> round kernel chained programmatically; GPU suite, bench (20 / 448 steps), a one-GPU C5 rehearsal
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2i1_pytest_gpu.txt; tail -5 gpurun_out/r2i1_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = b['roofline']; t = r.get('timeline') or {}
print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e (%s) launches %s parity %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b['e2e'].get('api'), b.get('gpu_launches'), b.get('parity_check')))
print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'), b['e2e'].get('notes'))
print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
print('    split', {k: round(v, 2) for k, v in r['split_kernels_us'].items()}, 'frac %.3f' % r['frac'])
g = b.get('state_machine_workload')
if g: print('    ring: value %.3e us/round %.2f conv %s applied/sent %.3f parity %s' % (g['value'], g['ms_per_step']*1e3, g['rounds_to_convergence'], g['recs_applied_over_recs_sent'], g['parity_check']), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (g.get('timeline') or {}).items() if k != 'what'})
PY
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2i1_bench20.json 2> gpurun_out/r2i1_bench20.err; tail -2 gpurun_out/r2i1_bench20.err | cut -c1-200; show gpurun_out/r2i1_bench20.json
timeout 300 python bench.py --no-cpu > gpurun_out/r2i1_bench448.json 2> gpurun_out/r2i1_bench448.err; tail -2 gpurun_out/r2i1_bench448.err | cut -c1-200; show gpurun_out/r2i1_bench448.json
timeout 200 python studies/c5_suspicion_sweep.py --nodes-per-gpu 2097152 --rounds 300 --suspicion 3 --sample-every 100 > gpurun_out/r2i1_c5.jsonl 2> gpurun_out/r2i1_c5.err
tail -2 gpurun_out/r2i1_c5.err | cut -c1-300
grep "^{" gpurun_out/r2i1_c5.jsonl | python -c "
import json,sys
for l in sys.stdin:
    b=json.loads(l); print('C5 1 GPU S', b['config']['S'], 'device us/round', b.get('device_us_per_round_rank0'), 'wall', b['wall_s'], b['detection_latency_rounds'], b['setup_s'])
"
