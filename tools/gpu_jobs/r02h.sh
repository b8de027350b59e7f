#!/bin/bash
# Round 2, job h: stream priorities with the H chain split into NTT pipeline (s[0]) and hExps MSM (s[5]).
# ZKR_STREAM_PRIO order: H-ntt, A, B1, B2, C, H-msm, upload.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-cpu --no-batch-2p22 --steps 10 > gpurun_out/r02h_$name.json 2>gpurun_out/r02h_$name.err || tail -3 gpurun_out/r02h_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02h_$name.json").read().strip().splitlines()[-1])
    print("%-12s" % "$name", d["ms_per_step"], d["e2e"]["ms_per_step"], d["prove_ms_serial"], d["stage_ms_overlapped"])
except Exception as e:
    print("$name failed", e)
PY
}
run base ZKR_X=0
run split ZKR_H_SPLIT=1
run p_ntt ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-1,0,0,0,0,0,0
run p_ntt_ab ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,-1,-1,0,0,0,0
run p_ntt_abb2 ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,-1,-1,-1,0,0,0
run p_ntt_b2 ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,0,0,-1,0,0,0
run p_all_but_hm ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-1,-1,-1,-1,-1,0,0
health end
