#!/bin/bash
# Round 2: the bench exactly as the driver launches it at N = 8 (torchrun, NCCL, two real GPUs, peer memory over NVLink).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo rc=$?
grep "\[bench\]" gpurun_out/r02_bench_n8.err | tail -12
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
    s = d["sharded"]; print("sharded identical", s["identical"], "proof", s["proof"]["ms"], s["proof"]["single_gpu_ms"], s["proof"]["identical"])
    for r in s["ntt"]: print("ntt", r["log_n"], r["dif_ms"], r["single_gpu_dif_ms"], r["speedup_vs_1gpu"], r["nvlink_gb_per_s_per_rank"], r["identical"])
    for r in s["msm_g1"]: print("msm", r["log_n"], r["ms"], r["single_gpu_ms"], r["speedup_vs_1gpu"], r["identical"])
    b = d["batch_2p22"]; print("batch", b["proofs"], b["ms"], b["proofs_per_s"], b["identical"])
except Exception as e:
    print("parse failed", e)
PY
timeout 30 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
