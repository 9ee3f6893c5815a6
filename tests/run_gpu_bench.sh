#!/bin/bash
# final single-GPU pass: GPU tests, smoke, bench (+ CPU arm), ncu launch list and --set full captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.txt
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 64 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python tests/prof_target.py 40 > gpurun_out/ncu_launch.log 2>&1
SWIM_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_split.csv \
    python tests/prof_target.py 40 > gpurun_out/ncu_launch2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 5 -c 1 -f -o gpurun_out/prof_round_quiet \
    python tests/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 27 -c 1 -f -o gpurun_out/prof_round_burst \
    python tests/prof_target.py 30 > gpurun_out/ncu_full1.log 2>&1
SWIM_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:tick_scan_kernel -s 5 -c 1 -f -o gpurun_out/prof_scan_quiet \
    python tests/prof_target.py 8 > gpurun_out/ncu_full2.log 2>&1
SWIM_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:'tick_|recv_' -s 75 -c 3 -f -o gpurun_out/prof_burst \
    python tests/prof_target.py 28 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | head -40
