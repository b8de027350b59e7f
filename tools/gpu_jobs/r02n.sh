#!/bin/bash
# Round 2, job n: window size c with the round-2 tail (gather) -- round 1 found c = 17 best with the recursive levels.
set -u
mkdir -p gpurun_out
for c in 17 18 19 20 21; do
  ZKR_MSM_C=$c timeout 400 python bench.py --no-cpu --no-batch-2p22 --no-gpu-witness --steps 12 > gpurun_out/r02n_c$c.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02n_c$c.json").read().strip().splitlines()[-1]); e = d["e2e"]
print($c, d["ms_per_step"], e["one_in_flight"]["ms_per_step"], e["two_in_flight"]["ms_per_step"], d["prove_ms_serial"], d["roofline"]["avg_launch_ms"], d["config"]["l2"][:40])
PY
done
timeout 30 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
