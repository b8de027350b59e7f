#!/bin/bash
# Round 2, job m: 256-thread bucket-reduction CTAs for small MSMs: parity, then A/B on the small circuits and small sweeps.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest (MSM, proofs, KATs, sharded, verify)"
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_golden_kats.py tests/test_gpu_sharded.py tests/test_verify.py -m gpu -x -q 2>&1 | tail -4
health after-tests
for shape in tx withdraw; do
  for w in 1 0 1 0; do
    ZKR_REDUCE_WIDE=$w timeout 300 python bench.py --shape $shape --no-cpu --no-batch-2p22 --no-gpu-witness --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$shape wide=$w', d['ms_per_step'], e['one_in_flight']['ms_per_step'], e['two_in_flight']['ms_per_step'], d['prove_ms_serial'], d['stage_ms_overlapped']['msm_b2_ms'])"
  done
done
echo "== standalone MSM 2^16 / 2^18"
for w in 1 0; do ZKR_REDUCE_WIDE=$w timeout 300 python tools/sweep.py --min-log 16 --max-log 18 --g2-min-log 16 --g2-max-log 18 --skip-ntt --out gpurun_out/r02m_sweep_w$w.json | grep uniform | cut -c1-140 | sed "s/^/wide=$w /"; done
echo "== 2^20 default (must not change)"
timeout 300 python bench.py --no-cpu --no-batch-2p22 --no-gpu-witness --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['two_in_flight'], d['prove_ms_serial'])"
health end
