#!/bin/bash
# Round 2, call E1 (1 GPU): GPU suite (churn, shards on one device with the new handshake), bench, calibration,
# ncu --set full of the burst launch of round_kernel, HBM-regime point (2^24 nodes on one GPU) with its ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2e_pytest_gpu.txt; tail -5 gpurun_out/r2e_pytest_gpu.txt
python - <<'PY' 2>&1 | tail -2
import json
from swim_b200.sim import Simulator, default_config, generate_topology
sim = Simulator(default_config(n_nodes=1 << 20, device=0))
sim.set_view(generate_topology("random", 1 << 20, 32, 32, seed=3))
sim.calibrate()
c = sim.calibrate()
print(c)
json.dump(c, open("gpurun_out/r2e_calibration.json", "w"))
PY
cp gpurun_out/r2e_calibration.json profiles/calibration.json
show() { python - "$1" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = b['roofline']; t = r.get('timeline') or {}
print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s parity %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches'), b.get('parity_check')))
print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'))
print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
print('    split', {k: round(v, 2) for k, v in r['split_kernels_us'].items()}, 'frac %.3f' % r['frac'], 'floor', r.get('latency_floor'))
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench20.json 2> gpurun_out/r2e_bench20.err; tail -3 gpurun_out/r2e_bench20.err; show gpurun_out/r2e_bench20.json
timeout 600 python bench.py --no-cpu > gpurun_out/r2e_bench448.json 2> gpurun_out/r2e_bench448.err; show gpurun_out/r2e_bench448.json
# launch list (cold, serialised) and the full capture of the burst launch (rounds 10..25 = the third round_kernel launch)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2e_launches20.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --windows 1 --spinup 0 > gpurun_out/r2e_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 2 -c 1 -f -o gpurun_out/r2e_round_kernel_burst python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --windows 1 --spinup 0 > gpurun_out/r2e_ncu2.log 2>&1
ncu -i gpurun_out/r2e_round_kernel_burst.ncu-rep --page raw --csv > gpurun_out/r2e_round_kernel_burst.csv 2>/dev/null
python tests/ncu_summary.py gpurun_out/r2e_round_kernel_burst.csv > gpurun_out/r2e_round_kernel_burst.txt 2>&1; head -8 gpurun_out/r2e_round_kernel_burst.txt
# HBM regime: 2^24 nodes on one GPU
timeout 900 python tests/prof_hbm_regime.py 24 gpurun_out/r2e_hbm_regime.json 2>&1 | tail -2
SWIM_SPLIT=1 timeout 900 ncu --set full --clock-control none -k regex:tick_scan -s 20 -c 2 -f -o gpurun_out/r2e_hbm_scan python tests/prof_hbm_regime.py 24 > gpurun_out/r2e_ncu3.log 2>&1
ncu -i gpurun_out/r2e_hbm_scan.ncu-rep --page raw --csv > gpurun_out/r2e_hbm_scan.csv 2>/dev/null
python tests/ncu_summary.py gpurun_out/r2e_hbm_scan.csv > gpurun_out/r2e_hbm_scan.txt 2>&1; head -12 gpurun_out/r2e_hbm_scan.txt
ls -la gpurun_out/*.ncu-rep
