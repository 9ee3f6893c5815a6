#!/bin/bash
# bench + launch list + one full ncu capture of the tick kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.txt
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 64 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python tests/prof_target.py 40 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tick_scan_kernel -s 5 -c 1 -f -o gpurun_out/prof_scan_quiet \
    python tests/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tick_|recv_' -s 75 -c 3 -f -o gpurun_out/prof_burst \
    python tests/prof_target.py 28 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
