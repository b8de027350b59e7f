#!/bin/bash
# Round-1 job H: evidence for round 2 — NTT pass kernels (DIF vs DIT) and the G2 accumulation under ncu --set full with
# source pages, plus NTT-only timings per variant.
set -u
mkdir -p gpurun_out
echo "== ntt probe timings"; for lg in 20 22 24; do timeout 200 python tools/ntt_probe.py --log-n $lg --reps 20 2>&1 | tail -6; done | tee gpurun_out/r01h_ntt_probe.jsonl
echo "== ncu full: ntt passes (2^20)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 12 -f -o gpurun_out/r01h_ntt python tools/ntt_probe.py --log-n 20 --reps 1 --no-time > gpurun_out/r01h_ncu_ntt.log 2>&1; tail -c 200 gpurun_out/r01h_ncu_ntt.log
ncu -i gpurun_out/r01h_ntt.ncu-rep --page raw --csv > gpurun_out/r01h_ntt_raw.csv 2>/dev/null
ncu -i gpurun_out/r01h_ntt.ncu-rep --page source --csv --print-source sass > gpurun_out/r01h_ntt_source.csv 2>/dev/null
echo "== ncu full: G2 accumulate"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_accum_affine.*Fq2' -c 1 -f -o gpurun_out/r01h_g2 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r01h_ncu_g2.log 2>&1; tail -c 200 gpurun_out/r01h_ncu_g2.log
ncu -i gpurun_out/r01h_g2.ncu-rep --page raw --csv > gpurun_out/r01h_g2_raw.csv 2>/dev/null
ncu -i gpurun_out/r01h_g2.ncu-rep --page source --csv --print-source sass > gpurun_out/r01h_g2_source.csv 2>/dev/null
ls -la gpurun_out; du -sm gpurun_out
if [ "$(du -sm gpurun_out | cut -f1)" -gt 60 ]; then rm -f gpurun_out/r01h_g2.ncu-rep; fi
if [ "$(du -sm gpurun_out | cut -f1)" -gt 60 ]; then rm -f gpurun_out/r01h_ntt.ncu-rep; fi
du -sm gpurun_out
