#!/bin/bash
# Round-1 measurement job B: config-5 unit (2^22 proof), fresh ncu launch list + full captures, big-size sweeps.
# Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
echo "== gpu tests (new json test only)"; timeout 300 python -m pytest tests/test_gpu_prove.py -m gpu -x -q -k "json" 2>&1 | tail -3
echo "== bench tx_2p22"; timeout 900 python bench.py --shape tx_2p22 --steps 5 --warmup 3 > gpurun_out/bench_2p22.json 2> gpurun_out/bench_2p22.err; cat gpurun_out/bench_2p22.json; tail -5 gpurun_out/bench_2p22.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1; tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-300
echo "== ncu full: accum"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accum_affine -s 5 -c 3 -f -o gpurun_out/r01b_accum python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_accum.log 2>&1; tail -1 gpurun_out/ncu_accum.log | cut -c1-200
echo "== ncu full: ntt + xyzz + bucket"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ntt_pass|k_accum_xyzz|k_bucket_sums|k_bucket_weighted' -s 60 -c 12 -f -o gpurun_out/r01b_rest python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rest.log 2>&1; tail -1 gpurun_out/ncu_rest.log | cut -c1-200
echo "== sweep G1 2^26, G2 2^24"; timeout 900 python tools/sweep.py --min-log 26 --max-log 26 --g2-min-log 24 --g2-max-log 24 --skip-ntt --reps 3 --out gpurun_out/sweep_big.json 2>&1 | tail -6
ls -la gpurun_out
