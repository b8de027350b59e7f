#!/bin/bash
set -u
mkdir -p gpurun_out
for k in 2 3 4; do
  timeout 400 python bench.py --no-cpu --no-batch-2p22 --no-gpu-witness --steps 12 --in-flight $k > gpurun_out/r02l_k$k.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02l_k$k.json").read().strip().splitlines()[-1]); e = d["e2e"]
print($k, d["ms_per_step"], e["one_in_flight"], e["two_in_flight"])
PY
done
timeout 30 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
