#!/bin/bash
# Round 2, FIRST gpurun call (1 GPU, ~12 min of box time): everything written after the last hardware run (R2_PREP_NOTES.md)
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tests/run_gpu_r2_first.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader | head -2
# 1. the whole GPU suite, most informative files first, no -x: one broken new test must not hide the rest
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.txt; tail -6 gpurun_out/r2_pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r2_smoke.txt
# 2. bench, default build (value must be within noise of round 1: 5.0-5.3e10; e2e > 1e10; convergence 306 / ~75 round-robin)
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
b = json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
r = b['roofline']
print('value %.3e  us/round %.2f  e2e %.3e  conv %s' % (b['value'], b['ms_per_step'] * 1e3, b['e2e']['value'], b['convergence']))
print({k: round(r[k] * 1e3, 2) for k in r if k.endswith('ms_per_launch')}, 'clocks', b.get('clocks'))
PY
# 2b. batched quiet scans A/B (R2_PREP_NOTES.md #11): off and 8 against the default (4) of step 2
for Q in 0 8; do
  SWIM_QUIET_BATCH=$Q timeout 300 python bench.py --no-cpu > gpurun_out/r2_bench_qb$Q.json 2> gpurun_out/r2_bench_qb$Q.err
  python - "$Q" <<'PY'
import json, sys
b = json.loads(open('gpurun_out/r2_bench_qb%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('SWIM_QUIET_BATCH', sys.argv[1], 'value %.3e  us/round %.2f' % (b['value'], b['ms_per_step'] * 1e3))
PY
done
# 3. CTA-size A/B (R2_PREP_NOTES.md #5): rebuild on the box, bench without the CPU arm, restore the default build
for W in 16 32; do
  SWIM_WPB=$W python -m swim_b200.build > /dev/null 2> gpurun_out/r2_build_wpb$W.err
  timeout 300 python bench.py --no-cpu > gpurun_out/r2_bench_wpb$W.json 2> gpurun_out/r2_bench_wpb$W.err
  python - "$W" <<'PY'
import json, sys
b = json.loads(open('gpurun_out/r2_bench_wpb%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('WPB', sys.argv[1], 'value %.3e  us/round %.2f' % (b['value'], b['ms_per_step'] * 1e3))
PY
done
python -m swim_b200.build --force > /dev/null
# 4. the C5 study at 64 Ki nodes, reference probe order and round-robin
timeout 300 python studies/c5_suspicion_sweep.py --nodes-per-gpu 65536 --rounds 300 --suspicion 2 5 13 > gpurun_out/r2_c5.jsonl 2> gpurun_out/r2_c5.err
timeout 300 python studies/c5_suspicion_sweep.py --nodes-per-gpu 65536 --rounds 300 --suspicion 2 5 13 --flags 2 > gpurun_out/r2_c5_rr.jsonl 2>> gpurun_out/r2_c5.err
cut -c1-400 gpurun_out/r2_c5.jsonl gpurun_out/r2_c5_rr.jsonl; tail -3 gpurun_out/r2_c5.err
ls -la gpurun_out | head -30
