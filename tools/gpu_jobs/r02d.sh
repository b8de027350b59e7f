#!/bin/bash
# Round 2, job d: parity of the reworked kernels (bucket gather, shared sort, fused H pipeline, full twiddle tables),
# then A/B timings against the round-1 paths (env knobs).
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
health after-tests
echo "== NTT probe: full twiddle tables vs two-level"
for lg in 20 22; do
  timeout 120 python tools/ntt_probe.py --log-n $lg --reps 20 | sed 's/^/twfull  /'
  ZKR_NTT_TWFULL_MAXLOG=0 timeout 120 python tools/ntt_probe.py --log-n $lg --reps 20 | sed 's/^/twolevel /'
done
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-cpu --no-batch-2p22 --steps 10 > gpurun_out/r02d_$name.json 2>gpurun_out/r02d_$name.err || tail -3 gpurun_out/r02d_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02d_$name.json").read().strip().splitlines()[-1])
    print("%-14s" % "$name", d["ms_per_step"], d["e2e"]["ms_per_step"], d["prove_ms_serial"], d["gpu_launches"], d["stage_ms_overlapped"], d["roofline"]["avg_launch_ms"], d["roofline_ntt"]["avg_launch_ms"])
except Exception as e:
    print("$name failed", e)
PY
}
run new ZKR_X=0
run levels ZKR_MSM_LEVELS=1
run hunfused ZKR_H_UNFUSED=1 ZKR_NTT_TWFULL_MAXLOG=0
run new2 ZKR_X=0
health end
