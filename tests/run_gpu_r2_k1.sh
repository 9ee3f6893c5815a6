#!/bin/bash
# Round 2, final single-GPU call K1: the records DESIGN.md section 8 / 11 cite (profiles/r02_final_*), most important first
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2k1_pytest_gpu.txt; tail -3 gpurun_out/r2k1_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json, sys
try:
    b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = b['roofline']; t = r.get('timeline') or {}
    print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s parity %s kernel %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches'), b.get('parity_check'), r.get('kernel')))
    print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'), b['e2e'].get('notes'))
    print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what' and v is not None})
    print('    frac %.3f dram_frac %s floor %s' % (r['frac'], r.get('dram_frac'), (r.get('latency_floor') or {}).get('frac_of_floor')))
    g = b.get('state_machine_workload')
    if g: print('    ring: value %.3e us/round %.2f conv %s applied/sent %.3f parity %s' % (g['value'], g['ms_per_step']*1e3, g['rounds_to_convergence'], g['recs_applied_over_recs_sent'], g['parity_check']))
    print('    cpu', b.get('cpu_baseline'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k1_bench20.json 2> gpurun_out/r2k1_bench20.err; tail -1 gpurun_out/r2k1_bench20.err | cut -c1-160; show gpurun_out/r2k1_bench20.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2k1_launches20.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --no-ring --windows 1 --spinup 0 > gpurun_out/r2k1_ncu1.log 2>&1
grep -c . gpurun_out/r2k1_launches20.csv
timeout 240 ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 2 -c 1 -f -o gpurun_out/r2k1_round_kernel_burst python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --no-ring --windows 1 --spinup 0 > gpurun_out/r2k1_ncu2.log 2>&1
ncu -i gpurun_out/r2k1_round_kernel_burst.ncu-rep --page raw --csv > gpurun_out/r2k1_round_kernel_burst.csv 2>/dev/null
python tests/ncu_summary.py gpurun_out/r2k1_round_kernel_burst.csv > gpurun_out/r2k1_round_kernel_burst.txt 2>&1; head -8 gpurun_out/r2k1_round_kernel_burst.txt
ncu -i gpurun_out/r2k1_round_kernel_burst.ncu-rep --page source --csv > gpurun_out/r2k1_round_kernel_burst_source.csv 2>/dev/null
timeout 200 python bench.py > gpurun_out/r2k1_bench448.json 2> gpurun_out/r2k1_bench448.err; tail -1 gpurun_out/r2k1_bench448.err | cut -c1-160; show gpurun_out/r2k1_bench448.json
timeout 100 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k1_ref20.json 2>/dev/null; cut -c1-200 gpurun_out/r2k1_ref20.json
