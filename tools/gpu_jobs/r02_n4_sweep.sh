#!/bin/bash
# Round 2: configs 2-3 at 4 GPUs for every size: sharded G1 MSM 2^16 .. 2^24 and four-step NTT 2^16 .. 2^26 (bench.py's sharded block
# with explicit size lists; replica timing shortened, other blocks off).
set -u
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-batch-2p22 --no-gpu-witness --no-two-in-flight \
    --sharded-ntt-logs 16,18,20,22,24,26 --sharded-msm-log 16,18,20,22,24 > gpurun_out/r02_sharded_sweep_n4.json 2> gpurun_out/r02_sharded_sweep_n4.err; echo rc=$?
grep "\[bench\] sharded" gpurun_out/r02_sharded_sweep_n4.err
timeout 30 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -3
