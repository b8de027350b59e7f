#!/bin/bash
# Round-1 job G: full GPU suite on a fresh box (as the driver runs it), smoke, default bench, reference arm, fresh launch list.
set -u
mkdir -p gpurun_out
echo "== gpu tests (full)"; timeout 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -22
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py > gpurun_out/r01g_bench_n1.json 2> gpurun_out/r01g_bench_n1.err; tail -c 1500 gpurun_out/r01g_bench_n1.json; tail -3 gpurun_out/r01g_bench_n1.err
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r01g_ncu_launch_bench.log 2>&1; tail -c 200 gpurun_out/r01g_ncu_launch_bench.log
