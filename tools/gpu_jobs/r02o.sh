#!/bin/bash
# Round 2: task-aware sharded proof (H group + weighted point ranges) -- emulated-rank parity on one GPU, both plans.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== sharded tests, task plan (default)"
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -4
health tasks
echo "== sharded tests, uniform plan"
ZKR_SHARD_TASKS=0 timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k prove 2>&1 | tail -4
health uniform
echo "== prove tests (one GPU path untouched)"
timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -x -q -k "not full_size and not 2p22" 2>&1 | tail -4
health end
