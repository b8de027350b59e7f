#!/bin/bash
# Round-1 measurement job C: new GPU tests (verify, json), smoke, fresh ncu launch list + full captures (size-bounded).
set -u
mkdir -p gpurun_out
echo "== gpu tests: verify + json + api"; timeout 600 python -m pytest tests/test_verify.py tests/test_gpu_prove.py -m gpu -x -q -k "verify or json or api or setup" --durations=5 2>&1 | tail -12
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1; tail -c 300 gpurun_out/ncu_launch_bench.log
echo "== ncu full: accum"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accum_affine -s 5 -c 2 -f -o gpurun_out/r01c_accum python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_accum.log 2>&1; tail -c 200 gpurun_out/ncu_accum.log
echo "== ncu full: ntt + xyzz + bucket"; timeout 600 ncu --set full --clock-control none -k regex:'k_ntt_pass|k_accum_xyzz|k_bucket_sums|k_bucket_weighted|k_sparse_lc|k_blind' -s 70 -c 14 -f -o gpurun_out/r01c_rest python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_rest.log 2>&1; tail -c 200 gpurun_out/ncu_rest.log
for f in r01c_accum r01c_rest; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out; du -sm gpurun_out
# keep the merge under the 64 MiB cap: drop the larger report if needed (its raw CSV stays)
if [ "$(du -sm gpurun_out | cut -f1)" -gt 60 ]; then rm -f gpurun_out/r01c_rest.ncu-rep; fi
if [ "$(du -sm gpurun_out | cut -f1)" -gt 60 ]; then rm -f gpurun_out/r01c_accum.ncu-rep; fi
du -sm gpurun_out
