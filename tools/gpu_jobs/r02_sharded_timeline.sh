#!/bin/bash
# Round 2: where the sharded proof's time goes at N ranks (ZKR_TIMELINE=1 prints every stage event's offset from the start
# of the call): task-aware split (H group + weighted ranges, default) vs uniform split with the H pipeline on every rank;
# ALL=1 adds blinding after the gather and H-chain-first.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
N=${N:-4}
run() {
    tag=$1; shift
    env "$@" ZKR_TIMELINE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29547 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-batch-2p22 --no-gpu-witness \
        --no-two-in-flight --sharded-ntt-logs "" --sharded-msm-log "" \
        > gpurun_out/r02_tl_${tag}_n$N.json 2> gpurun_out/r02_tl_${tag}_n$N.err
    echo "== $tag rc=$?"
    grep "zkr timeline rank [0-9]/$N" gpurun_out/r02_tl_${tag}_n$N.err | tail -$((3 * N))
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_tl_${tag}_n$N.json").read().strip().splitlines()[-1])
    p = d["sharded"]["proof"]; print("$tag", p["ms"], p["single_gpu_ms"], p["identical"], p["stage_ms"])
except Exception as e:
    print("parse failed", e)
PY
}
run tasks X=0
run uniform ZKR_SHARD_TASKS=0
if [ "${ALL:-0}" = 1 ]; then
    run late ZKR_SHARDED_BLIND_LATE=1
    run hfirst ZKR_H_FIRST=1
fi
timeout 30 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
