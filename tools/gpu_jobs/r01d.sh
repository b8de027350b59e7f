#!/bin/bash
# job D: correctness of the boundary-level change, then the scheduling / level-granularity sweep
set -u
mkdir -p gpurun_out
echo "== gpu tests: msm + prove(small)"; timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -m gpu -x -q -k "not full_size" 2>&1 | tail -4
echo "== sweep"; timeout 900 python tools/prio_sweep.py --steps 10 2>&1 | tail -14
