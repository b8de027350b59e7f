#!/bin/bash
# Round 2, job r: dedicated squaring inside the lazily reduced G1 addition (fp.cuh mont_sqr_raw: 100 instead of 128 wide
# multiplies per squaring; ZKR_LAZY bit 2).  Device parity first, then the A/B against the current default (ZKR_LAZY=3).
set -u
mkdir -p gpurun_out
echo "== field / curve / msm parity (all accumulation forms)"
timeout 600 python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py -m gpu -x -q -k "not full_size" 2>&1 | tail -4
echo "== proofs with ZKR_LAZY=7"
ZKR_LAZY=7 timeout 400 python -m pytest tests/test_golden_kats.py tests/test_gpu_prove.py -m gpu -x -q -k "golden or bit_exact_small or invalid_witness" 2>&1 | tail -3
echo "== microbench"
timeout 200 python tools/microbench.py gpurun_out/r02r_microbench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: (round(v/1e9,3) if k.endswith('_per_s') else round(v,3)) for k,v in d.items() if 'g1_madd' in k or 'modmul_per' in k})"
run() {
    tag=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --no-batch-2p22 --no-gpu-witness --steps 12 > gpurun_out/r02r_$tag.json 2>gpurun_out/r02r_$tag.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02r_$tag.json").read().strip().splitlines()[-1]); e = d["e2e"]
    print("$tag", "dev", d["ms_per_step"], "e2e1", e["one_in_flight"]["ms_per_step"], "e2e2", e["two_in_flight"]["ms_per_step"], "serial", d["prove_ms_serial"],
          "g1_launch", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"])
except Exception as ex:
    print("$tag failed", ex)
PY
}
run lazy3_a ZKR_LAZY=3
run lazy7_a ZKR_LAZY=7
run lazy3_b ZKR_LAZY=3
run lazy7_b ZKR_LAZY=7
for v in "lazy3 ZKR_LAZY=3" "lazy7 ZKR_LAZY=7"; do
    set -- $v; tag=$1; shift
    env "$@" timeout 200 python bench.py --shape tx --no-cpu --no-batch-2p22 --no-gpu-witness --steps 20 > gpurun_out/r02r_tx_$tag.json 2>/dev/null
    python -c "
import json; d=json.loads(open('gpurun_out/r02r_tx_$tag.json').read().strip().splitlines()[-1]); e=d['e2e']
print('tx $tag', d['ms_per_step'], e['one_in_flight']['ms_per_step'], e['two_in_flight']['ms_per_step'])"
done
