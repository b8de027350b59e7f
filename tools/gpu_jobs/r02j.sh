#!/bin/bash
# Round 2, job j: G2 accumulation with one point addition per lane pair: parity, then A/B against the one-thread kernel.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest (G2 MSM, proofs, KATs, sharded)"
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_golden_kats.py tests/test_gpu_sharded.py tests/test_verify.py -m gpu -x -q 2>&1 | tail -6
health after-tests
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --no-batch-2p22 --no-gpu-witness --steps 10 > gpurun_out/r02j_$name.json 2>gpurun_out/r02j_$name.err || tail -3 gpurun_out/r02j_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02j_$name.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("%-8s" % "$name", d["ms_per_step"], e["one_in_flight"], e["two_in_flight"], d["prove_ms_serial"], d["stage_ms_overlapped"])
except Exception as e:
    print("$name failed", e)
PY
}
run pair ZKR_G2_PAIR=1
run single ZKR_G2_PAIR=0
run pair2 ZKR_G2_PAIR=1
run single2 ZKR_G2_PAIR=0
echo "== standalone G2 MSM"
ZKR_G2_PAIR=1 timeout 600 python tools/sweep.py --min-log 30 --max-log 30 --g2-min-log 18 --g2-max-log 22 --skip-ntt --out gpurun_out/r02j_g2_pair.json | grep uniform | cut -c1-150 | sed 's/^/pair   /'
ZKR_G2_PAIR=0 timeout 600 python tools/sweep.py --min-log 30 --max-log 30 --g2-min-log 18 --g2-max-log 22 --skip-ntt --out gpurun_out/r02j_g2_single.json | grep uniform | cut -c1-150 | sed 's/^/single /'
health end
