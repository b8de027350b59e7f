#!/bin/bash
# Round 2, job b: new GPU tests (2^20 AP MSM, NTT Horner 2^20/2^24, tx_2p22, published vectors), bench N=1 with the
# batch_2p22 block, and the N=2 bench flow on one device (ZKR_BENCH_ONE_DEVICE=1: both ranks on cuda:0, gloo).
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (new tests first)"
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 -k "published or full_size or horner or geometric or arithmetic_progression" 2>&1 | tail -25
echo "== bench N=1"
timeout 600 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo rc=$?; tail -5 gpurun_out/r02b_bench_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02b_bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "prove_ms_serial")}, d["e2e"]["value"], d["batch_2p22"])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== bench N=2 flow test on one device"
ZKR_BENCH_ONE_DEVICE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --sharded-ntt-logs 22,24 --sharded-msm-log 22 > gpurun_out/r02b_bench_n2_onedev.json 2> gpurun_out/r02b_bench_n2_onedev.err; echo rc=$?
tail -12 gpurun_out/r02b_bench_n2_onedev.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02b_bench_n2_onedev.json").read().strip().splitlines()[-1])
    print(json.dumps(d["sharded"], indent=1)[:3000]); print(d["batch_2p22"])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== remaining GPU tests"
timeout 900 python -m pytest tests -m gpu -x -q -k "not (published or full_size or horner or geometric or arithmetic_progression)" 2>&1 | tail -6
