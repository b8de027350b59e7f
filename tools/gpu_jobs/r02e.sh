#!/bin/bash
# Round 2, job e: schedule knobs on top of the gather, the 1-GPU sweep with CPU baselines, CPU-A / CPU-B records at the
# tx.circom size, ncu captures of the reworked kernels.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-cpu --no-batch-2p22 --steps 10 > gpurun_out/r02e_$name.json 2>gpurun_out/r02e_$name.err || tail -3 gpurun_out/r02e_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02e_$name.json").read().strip().splitlines()[-1])
    print("%-14s" % "$name", d["ms_per_step"], d["e2e"]["ms_per_step"], d["prove_ms_serial"], d["stage_ms_overlapped"])
except Exception as e:
    print("$name failed", e)
PY
}
run base ZKR_X=0
run abch ZKR_DELAY=abch
run ABCH ZKR_DELAY=ABCH
run abc ZKR_DELAY=abc
run h ZKR_DELAY=h
run prioH ZKR_STREAM_PRIO=-1,0,0,0,0,0
run prioHB2 ZKR_STREAM_PRIO=-1,0,0,-1,0,0
health after-sched
echo "== sweep (configs 2-3, 1 GPU, CPU baselines)"
timeout 1500 python bench.py --sweep msm,ntt --sweep-max-log 26 --sweep-out gpurun_out/r02e_sweep_1gpu.json > gpurun_out/r02e_sweep.log 2> gpurun_out/r02e_sweep.err; echo sweep rc=$?; tail -3 gpurun_out/r02e_sweep.err; grep -c correct gpurun_out/r02e_sweep.log; grep -c '"correct": false' gpurun_out/r02e_sweep.log
health after-sweep
echo "== CPU-A / CPU-B at the tx.circom size (config 1)"
timeout 900 python bench.py --impl reference --shape tx --ref-mode 0 --ref-threads 1 --steps 1 --warmup 0 > gpurun_out/r02e_ref_tx_cpuA.json 2>gpurun_out/r02e_ref_tx_cpuA.err; tail -1 gpurun_out/r02e_ref_tx_cpuA.json | cut -c1-300
timeout 300 python bench.py --impl reference --shape tx --ref-mode 1 --steps 2 --warmup 1 > gpurun_out/r02e_ref_tx_cpuB.json 2>/dev/null; tail -1 gpurun_out/r02e_ref_tx_cpuB.json | cut -c1-300
timeout 300 python bench.py --shape tx --no-batch-2p22 --steps 20 > gpurun_out/r02e_bench_tx.json 2>/dev/null; tail -1 gpurun_out/r02e_bench_tx.json | cut -c1-200
echo "== ncu: reworked kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 6 -f -o gpurun_out/r02e_ntt \
    python tools/ntt_probe.py --log-n 20 --reps 1 --no-time > gpurun_out/r02e_ncu_ntt.log 2>&1; tail -c 150 gpurun_out/r02e_ncu_ntt.log
ncu -i gpurun_out/r02e_ntt.ncu-rep --page raw --csv > gpurun_out/r02e_ntt_raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_ntt.ncu-rep --page source --csv --print-source sass > gpurun_out/r02e_ntt_source.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/r02e_ntt_source.csv --top 25 > gpurun_out/r02e_ntt_source_top.txt 2>&1
gzip -9 -f gpurun_out/r02e_ntt_source.csv; rm -f gpurun_out/r02e_ntt.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-batch-2p22 > gpurun_out/r02e_ncu_launch_bench.log 2>&1; tail -c 200 gpurun_out/r02e_ncu_launch_bench.log
health end
du -sm gpurun_out
