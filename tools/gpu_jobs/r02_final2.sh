#!/bin/bash
# Round 2: the default bench line (the record the driver reproduces), the new ABI-behaviour tests, NTT pass-split probe.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== new tests"
timeout 600 python -m pytest tests/test_gpu_prove.py tests/test_gpu_witness.py -m gpu -x -q -k "flags or blinding or cache or witness or persistent or solver or solvable or generator" 2>&1 | tail -5
echo "== bench (default flags)"
SECONDS=0
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$? wall=${SECONDS}s"
grep "\[bench\]" gpurun_out/r02_bench_n1.err | tail -8
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "prove_ms_serial", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["one_in_flight"], d["e2e"]["two_in_flight"])
print(d["roofline"]["frac"], d["roofline"]["practical_peak_frac"], d["roofline_ntt"]["transform"], d["roofline_ntt"]["h_pipeline"])
print(d["gpu_witness"]); print(d["batch_2p22"]["proofs_per_s"], d["cpu_baseline"]["seconds_per_proof"], d["clocks"])
PY
health after-bench
echo "== NTT pass split at 2^20 / 2^21 / 2^22"
for sp in "" "9,11" "11,9"; do ZKR_NTT_SPLIT=$sp timeout 120 python tools/ntt_probe.py --log-n 20 --reps 20 | grep "forward_dif\|inverse_dit" | sed "s/^/split[$sp] /"; done
for sp in "" "10,11" "11,10"; do ZKR_NTT_SPLIT=$sp timeout 120 python tools/ntt_probe.py --log-n 21 --reps 20 | grep "forward_dif\|inverse_dit" | sed "s/^/split[$sp] /"; done
health end
