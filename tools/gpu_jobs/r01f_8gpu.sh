#!/bin/bash
# job E (8 GPUs): sharded NTT / MSM / proof over real peer memory, then the independent-proofs bench at N = 8
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== sharded check (8 GPUs)"
timeout 500 $TR --master-port 29511 tools/multigpu_check.py --ntt-logs 24,26 --msm-logs 22,24 --prove-shapes tx_2p20 --out gpurun_out/multigpu_8.json 2>&1 | tail -25
echo "== bench --gpus 8"
timeout 400 $TR --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; cat gpurun_out/bench_n8.json | cut -c1-600; tail -3 gpurun_out/bench_n8.err
