#!/bin/bash
# Round 2, call J1 (1 GPU): round_kernel_x (one grid barrier per round) against the two-phase round_kernel — GPU suite with the
# new default, then bench 20 / 448 steps both ways (SWIM_XMODE=1 / 0); bench.py on ONE handle
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/r2j1_pytest_gpu.txt; tail -4 gpurun_out/r2j1_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json, sys
try:
    b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = b['roofline']; t = r.get('timeline') or {}
    print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s parity %s kernel %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches'), b.get('parity_check'), r.get('kernel')))
    print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'), b['e2e'].get('notes'))
    print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what' and v is not None})
    g = b.get('state_machine_workload')
    if g: print('    ring: value %.3e us/round %.2f conv %s parity %s' % (g['value'], g['ms_per_step']*1e3, g['rounds_to_convergence'], g['parity_check']), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (g.get('timeline') or {}).items() if k != 'what' and v is not None})
    print('    conv', b.get('convergence'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
for X in 1 0; do
  SWIM_XMODE=$X timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2j1_bench20_x$X.json 2> gpurun_out/r2j1_bench20_x$X.err; tail -1 gpurun_out/r2j1_bench20_x$X.err | cut -c1-200; show gpurun_out/r2j1_bench20_x$X.json
done
for X in 1 0; do
  SWIM_XMODE=$X timeout 200 python bench.py --no-cpu --converge-limit 400 > gpurun_out/r2j1_bench448_x$X.json 2> gpurun_out/r2j1_bench448_x$X.err; tail -1 gpurun_out/r2j1_bench448_x$X.err | cut -c1-200; show gpurun_out/r2j1_bench448_x$X.json
done
