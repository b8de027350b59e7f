#!/bin/bash
# Round 2, job c: the GPU tests job b did not reach, then A/B of the proof schedule (shared B1'/B2' sort, chains
# delayed until the G2 accumulation is done).
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu: published vectors, fill, full-size proofs (incl. tx_2p22)"
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 -k "published or full_size or geometric" 2>&1 | tail -16
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-cpu --no-batch-2p22 --steps 10 > gpurun_out/r02c_$name.json 2>/dev/null
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02c_$name.json").read().strip().splitlines()[-1])
    print("%-18s" % "$name", d["ms_per_step"], d["e2e"]["ms_per_step"], d["prove_ms_serial"], d["stage_ms_overlapped"])
except Exception as e:
    print("$name failed", e)
PY
}
run share1 ZKR_SHARE_SORT=1
run share0 ZKR_SHARE_SORT=0
run delayC ZKR_DELAY=C
run delayCH ZKR_DELAY=CH
run delayA ZKR_DELAY=A
run delayAC ZKR_DELAY=AC
run delayH ZKR_DELAY=H
run share1b ZKR_SHARE_SORT=1
