#!/bin/bash
# Round 2, call B (1 GPU): event kernel by node runs, pinned event staging, kernel preload, bench with device checkpoint +
# clock spin-up + median of 5 windows, phase timeline, latency calibration; CTA-size A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2b_pytest_gpu.txt; tail -4 gpurun_out/r2b_pytest_gpu.txt
python - <<'PY' 2>&1 | tail -3
import json
from swim_b200.sim import Simulator, default_config, generate_topology
sim = Simulator(default_config(n_nodes=1 << 20, device=0))
sim.set_view(generate_topology("random", 1 << 20, 32, 32, seed=3))
c = sim.calibrate()
c2 = sim.calibrate()
print(c, c2)
json.dump(c2, open("gpurun_out/r2b_calibration.json", "w"))
PY
cp gpurun_out/r2b_calibration.json profiles/calibration.json
show() { python - "$1" <<'PY'
import json, sys
b = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = b['roofline']; t = r.get('timeline') or {}
print(sys.argv[1], 'value %.3e  us/round %.2f  e2e %.3e launches %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b.get('gpu_launches')))
print('    windows', b['timing']['windows_ms'], 'e2e windows', b['e2e'].get('windows_ms'))
print('    timeline', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in t.items() if k != 'what'})
print('    split', {k: round(v, 2) for k, v in r['split_kernels_us'].items()}, 'frac %.3f' % r['frac'], 'floor', r.get('latency_floor'))
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench20.json 2> gpurun_out/r2b_bench20.err; tail -3 gpurun_out/r2b_bench20.err; show gpurun_out/r2b_bench20.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench20_2.json 2> gpurun_out/r2b_bench20_2.err; show gpurun_out/r2b_bench20_2.json
timeout 600 python bench.py --no-cpu > gpurun_out/r2b_bench448.json 2> gpurun_out/r2b_bench448.err; show gpurun_out/r2b_bench448.json
SWIM_QUIET_BATCH=0 timeout 600 python bench.py --no-cpu > gpurun_out/r2b_bench448_qb0.json 2> gpurun_out/r2b_bench448_qb0.err; show gpurun_out/r2b_bench448_qb0.json
for W in 16 32; do
  SWIM_WPB=$W python -m swim_b200.build > /dev/null 2> gpurun_out/r2b_build_wpb$W.err
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench20_wpb$W.json 2> gpurun_out/r2b_bench20_wpb$W.err; show gpurun_out/r2b_bench20_wpb$W.json
  timeout 300 python bench.py --no-cpu > gpurun_out/r2b_bench448_wpb$W.json 2> gpurun_out/r2b_bench448_wpb$W.err; show gpurun_out/r2b_bench448_wpb$W.json
done
python -m swim_b200.build --force > /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2b_launches20.csv python bench.py --steps 20 --warmup 5 --no-cpu --windows 1 --spinup 0 > gpurun_out/r2b_ncu.log 2>&1
grep -c . gpurun_out/r2b_launches20.csv
