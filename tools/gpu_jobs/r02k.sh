#!/bin/bash
# Round 2, job k: zero digits as skip entries (c - 1 key bits) vs their own sentinel key (c bits): alternating A/B.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --no-batch-2p22 --no-gpu-witness --steps 20 > gpurun_out/r02k_$name.json 2>gpurun_out/r02k_$name.err || tail -3 gpurun_out/r02k_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02k_$name.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("%-8s" % "$name", d["ms_per_step"], e["one_in_flight"], e["two_in_flight"], d["prove_ms_serial"], d["roofline"]["avg_launch_ms"])
except Exception as e:
    print("$name failed", e)
PY
}
run skip ZKR_ZERO_SENTINEL=0
run sent ZKR_ZERO_SENTINEL=1
run skip2 ZKR_ZERO_SENTINEL=0
run sent2 ZKR_ZERO_SENTINEL=1
run skip3 ZKR_ZERO_SENTINEL=0
run sent3 ZKR_ZERO_SENTINEL=1
echo "== quick parity with the sentinel layout"
ZKR_ZERO_SENTINEL=1 timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_golden_kats.py -m gpu -x -q 2>&1 | tail -3
health end
