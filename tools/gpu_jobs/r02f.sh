#!/bin/bash
# Round 2, job f: parity after the merged gather / radix-4 NTT stages / pipelined twiddles, then A/B.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
health after-tests
echo "== NTT probe: radix-8 vs radix-4 register stages"
for lg in 18 20 22; do
  ZKR_NTT_RADIX_LOG=3 timeout 120 python tools/ntt_probe.py --log-n $lg --reps 20 | grep "forward_dif\|inverse_dit" | sed 's/^/radix8 /'
  ZKR_NTT_RADIX_LOG=2 timeout 120 python tools/ntt_probe.py --log-n $lg --reps 20 | grep "forward_dif\|inverse_dit" | sed 's/^/radix4 /'
done
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-cpu --no-batch-2p22 --steps 10 > gpurun_out/r02f_$name.json 2>gpurun_out/r02f_$name.err || tail -3 gpurun_out/r02f_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02f_$name.json").read().strip().splitlines()[-1])
    print("%-10s" % "$name", d["ms_per_step"], d["e2e"]["ms_per_step"], d["prove_ms_serial"], d["gpu_launches"], d["stage_ms_overlapped"], d["roofline"]["avg_launch_ms"], d["roofline_ntt"]["avg_launch_ms"])
except Exception as e:
    print("$name failed", e)
PY
}
run radix8 ZKR_NTT_RADIX_LOG=3
run radix4 ZKR_NTT_RADIX_LOG=2
echo "== standalone G1 MSM: gather vs recursive levels"
timeout 600 python tools/sweep.py --min-log 18 --max-log 24 --g2-max-log 0 --skip-ntt --out gpurun_out/r02f_sweep_gather.json | grep uniform | cut -c1-160 | sed 's/^/gather /'
ZKR_MSM_LEVELS=1 timeout 600 python tools/sweep.py --min-log 18 --max-log 24 --g2-max-log 0 --skip-ntt --out gpurun_out/r02f_sweep_levels.json | grep uniform | cut -c1-160 | sed 's/^/levels /'
echo "== microbench"
timeout 300 python tools/microbench.py | tr ',' '\n' | grep "inverse\|batch\|madd\|modmul"
cp gpurun_out/microbench.json gpurun_out/r02f_microbench.json
health end
