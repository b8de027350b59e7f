#!/bin/bash
# Round 2, job h: parity after the skip-entry sort change; stream priorities with the H chain split into NTT pipeline
# (s[0]) and hExps MSM (s[5]) -- ZKR_STREAM_PRIO order: H-ntt, A, B1, B2, C, H-msm, upload; ncu --set full of the gather.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
health after-tests
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-cpu --no-batch-2p22 --steps 10 > gpurun_out/r02h_$name.json 2>gpurun_out/r02h_$name.err || tail -3 gpurun_out/r02h_$name.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02h_$name.json").read().strip().splitlines()[-1])
    print("%-12s" % "$name", d["ms_per_step"], d["e2e"]["ms_per_step"], d["prove_ms_serial"], d["stage_ms_overlapped"], d["roofline_ntt"].get("transform", {}).get("ms"), d["roofline_ntt"].get("h_pipeline", {}).get("ms"))
except Exception as e:
    print("$name failed", e)
PY
}
run base ZKR_X=0
run split ZKR_H_SPLIT=1
run p_ntt ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-1,0,0,0,0,0,0
run p_ntt_ab ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,-1,-1,0,0,0,0
run p_ntt_abb2 ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,-1,-1,-1,0,0,0
run p_ntt_b2 ZKR_H_SPLIT=1 ZKR_STREAM_PRIO=-2,0,0,-1,0,0,0
run base2 ZKR_X=0
health after-bench
echo "== ncu full: k_bucket_gather at 2^22 (c = 20) and 2^24 (c = 22, gather forced)"
for lg in 22 24; do
  ZKR_MSM_LEVELS=0 timeout 600 ncu --set full --clock-control none -k regex:k_bucket_gather -c 1 -f -o gpurun_out/r02h_gather$lg \
      python tools/sweep.py --min-log $lg --max-log $lg --g2-max-log 0 --skip-ntt --reps 1 --out gpurun_out/r02h_tmp.json > gpurun_out/r02h_ncu_gather$lg.log 2>&1
  ncu -i gpurun_out/r02h_gather$lg.ncu-rep --page raw --csv > gpurun_out/r02h_gather${lg}_raw.csv 2>/dev/null
  rm -f gpurun_out/r02h_gather$lg.ncu-rep
  python tools/ncu_summary.py raw gpurun_out/r02h_gather${lg}_raw.csv gpurun_out/r02h_gather${lg}.md > /dev/null 2>&1; grep -v "^$" gpurun_out/r02h_gather${lg}.md | head -60
done
health end
