#!/bin/bash
# Round 2, job i: parity after spreading the skip entries / faster heavy scan; bench with the two-in-flight and
# witness-on-GPU legs; default schedule vs (H split + priorities); gather at 2^22 / 2^24 again.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
health after-tests
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --no-batch-2p22 --steps 10 > gpurun_out/r02i_$name.json 2>gpurun_out/r02i_$name.err || tail -3 gpurun_out/r02i_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02i_$name.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("%-10s" % "$name", d["ms_per_step"], e["one_in_flight"], e["two_in_flight"], d["prove_ms_serial"], d["stage_ms_overlapped"])
    print("           gpu_witness", d["gpu_witness"])
except Exception as e:
    print("$name failed", e)
PY
}
run base ZKR_X=0
run prio ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,-1,-1,-1,0,0,0
run base2 ZKR_X=0
run prio2 ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,-1,-1,-1,0,0,0
health after-bench
echo "== standalone G1 MSM: gather (forced) vs levels (forced)"
ZKR_MSM_LEVELS=0 timeout 600 python tools/sweep.py --min-log 20 --max-log 24 --g2-max-log 0 --skip-ntt --out gpurun_out/r02i_sweep_gather.json | cut -c1-150 | sed 's/^/gather /'
ZKR_MSM_LEVELS=1 timeout 600 python tools/sweep.py --min-log 20 --max-log 24 --g2-max-log 0 --skip-ntt --out gpurun_out/r02i_sweep_levels.json | cut -c1-150 | sed 's/^/levels /'
health end
