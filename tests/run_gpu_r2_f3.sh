#!/bin/bash
# Round 2, call F3 (1 GPU): the final single-GPU build — GPU suite, bench (driver's command + long window), launch list,
# full ncu capture of the burst launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2f3_pytest_gpu.txt; tail -5 gpurun_out/r2f3_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = b['roofline']; t = r.get('timeline') or {}
print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e (%s) launches %s parity %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b['e2e'].get('api'), b.get('gpu_launches'), b.get('parity_check')))
print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'), b['e2e'].get('notes'))
print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
print('    split', {k: round(v, 2) for k, v in r['split_kernels_us'].items()}, 'frac %.3f' % r['frac'], 'floor', r.get('latency_floor'))
g = b.get('state_machine_workload')
if g: print('    ring: value %.3e us/round %.2f conv %s applied/sent %.3f parity %s' % (g['value'], g['ms_per_step']*1e3, g['rounds_to_convergence'], g['recs_applied_over_recs_sent'], g['parity_check']))
print('    cpu', b.get('cpu_baseline'))
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f3_bench20.json 2> gpurun_out/r2f3_bench20.err; tail -3 gpurun_out/r2f3_bench20.err; show gpurun_out/r2f3_bench20.json
timeout 600 python bench.py > gpurun_out/r2f3_bench448.json 2> gpurun_out/r2f3_bench448.err; tail -3 gpurun_out/r2f3_bench448.err; show gpurun_out/r2f3_bench448.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f3_ref20.json 2>/dev/null; cut -c1-300 gpurun_out/r2f3_ref20.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2f3_launches20.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --no-ring --windows 1 --spinup 0 > gpurun_out/r2f3_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 2 -c 1 -f -o gpurun_out/r2f3_round_kernel_burst python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --no-ring --windows 1 --spinup 0 > gpurun_out/r2f3_ncu2.log 2>&1
ncu -i gpurun_out/r2f3_round_kernel_burst.ncu-rep --page raw --csv > gpurun_out/r2f3_round_kernel_burst.csv 2>/dev/null
python tests/ncu_summary.py gpurun_out/r2f3_round_kernel_burst.csv > gpurun_out/r2f3_round_kernel_burst.txt 2>&1; head -8 gpurun_out/r2f3_round_kernel_burst.txt
