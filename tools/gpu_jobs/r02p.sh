#!/bin/bash
# Round 2, job p: lazily reduced additions (sums of products with one reduction: fp.cuh mont_mul2_raw / mont_mul4_raw,
# XYZZ::madd_lazy; ZKR_LAZY bit 0 = G1, bit 1 = G2 accumulation; libzkr_tail.so = the full XYZZ addition lazily reduced too).
# Device parity of the new arithmetic first, then the A/B.
set -u
mkdir -p gpurun_out
TAIL=$PWD/simple_zk_rollups_b200/libzkr_tail.so
echo "== field / curve / msm parity (both accumulation forms)"
timeout 600 python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py -m gpu -x -q -k "not full_size" 2>&1 | tail -4
echo "== proofs with ZKR_LAZY=3"
ZKR_LAZY=3 timeout 400 python -m pytest tests/test_golden_kats.py tests/test_gpu_prove.py -m gpu -x -q -k "golden or bit_exact_small or invalid_witness" 2>&1 | tail -3
echo "== tail library: field / curve / msm / proofs"
ZKR_LIB=$TAIL ZKR_LAZY=3 timeout 600 python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py tests/test_golden_kats.py tests/test_gpu_prove.py -m gpu -x -q -k "(field or curve or msm_small or golden or bit_exact_small) and not full_size" 2>&1 | tail -3
echo "== microbench"
timeout 200 python tools/microbench.py gpurun_out/r02p_microbench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: (round(v/1e9,3) if k.endswith('_per_s') else round(v,3)) for k,v in d.items() if 'madd' in k or 'modmul_per' in k})"
run() {
    tag=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --no-batch-2p22 --no-gpu-witness --steps 12 > gpurun_out/r02p_$tag.json 2>gpurun_out/r02p_$tag.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02p_$tag.json").read().strip().splitlines()[-1]); e = d["e2e"]
    print("$tag", "dev", d["ms_per_step"], "e2e1", e["one_in_flight"]["ms_per_step"], "e2e2", e["two_in_flight"]["ms_per_step"], "serial", d["prove_ms_serial"],
          "g1_launch", d["roofline"]["avg_launch_ms"], "b2", d["stage_ms_overlapped"]["msm_b2_ms"])
except Exception as ex:
    print("$tag failed", ex)
PY
}
run lazy0 ZKR_LAZY=0
run lazy1 ZKR_LAZY=1
run lazy2 ZKR_LAZY=2
run lazy3 ZKR_LAZY=3
run tail0 ZKR_LIB=$TAIL ZKR_LAZY=0
run tail3 ZKR_LIB=$TAIL ZKR_LAZY=3
for shape in tx withdraw; do
  for v in "lazy0 ZKR_LAZY=0" "lazy3 ZKR_LAZY=3" "tail3 ZKR_LAZY=3 ZKR_LIB=$TAIL"; do
    set -- $v; tag=$1; shift
    env "$@" timeout 200 python bench.py --shape $shape --no-cpu --no-batch-2p22 --no-gpu-witness --steps 20 > gpurun_out/r02p_${shape}_$tag.json 2>/dev/null
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02p_${shape}_$tag.json").read().strip().splitlines()[-1]); e = d["e2e"]
    print("$shape $tag", "dev", d["ms_per_step"], "e2e1", e["one_in_flight"]["ms_per_step"], "e2e2", e["two_in_flight"]["ms_per_step"])
except Exception as ex:
    print("$shape $tag failed", ex)
PY
  done
done
timeout 30 nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
