#!/bin/bash
# Round-1 job I (last of the round's GPU budget): validate the NTT load/twiddle restructure, the narrow-tile rule and the
# variable-time k_finish inversion with the full GPU suite, then bench + A/B timings, then (if time is left) the
# accumulation kernels under ncu with source pages.
set -u
T0=$(date +%s)
mkdir -p gpurun_out
echo "== gpu tests (full)"; timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench"; timeout 200 python bench.py > gpurun_out/r01i_bench_n1.json 2> gpurun_out/r01i_bench_n1.err; head -c 400 gpurun_out/r01i_bench_n1.json; echo; tail -2 gpurun_out/r01i_bench_n1.err
echo "== ntt probe: narrow tiles (default) vs full-width tiles"
for lg in 16 18 20 22; do
  timeout 60 python tools/ntt_probe.py --log-n $lg --reps 20 2>&1 | grep log_n | sed 's/^{/{"tiles": "narrow", /'
  ZKR_NTT_MIN_BLOCKS=0 timeout 60 python tools/ntt_probe.py --log-n $lg --reps 20 2>&1 | grep log_n | sed 's/^{/{"tiles": "full", /'
done | tee gpurun_out/r01i_ntt_probe.jsonl | cut -c1-150
echo "== small circuits: k_finish binary-Euclid (default) vs Fermat"
for shape in withdraw tx; do
  timeout 90 python bench.py --shape $shape --no-cpu --steps 20 > gpurun_out/r01i_bench_${shape}.json 2>/dev/null
  ZKR_FINISH_FERMAT=1 timeout 90 python bench.py --shape $shape --no-cpu --steps 20 > gpurun_out/r01i_bench_${shape}_fermat.json 2>/dev/null
  python - <<PY
import json
for f in ("gpurun_out/r01i_bench_${shape}.json", "gpurun_out/r01i_bench_${shape}_fermat.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("prove_ms_serial"))
    except Exception as e:
        print(f, "failed", e)
PY
done
NOW=$(date +%s); echo "elapsed $((NOW-T0)) s"
if [ $((NOW-T0)) -lt 470 ]; then
  echo "== ncu full: accumulation kernels (5 launches of one proof)"
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_accum_affine -s 5 -c 5 -f -o gpurun_out/r01i_accum python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r01i_ncu_accum.log 2>&1; tail -c 150 gpurun_out/r01i_ncu_accum.log
  ncu -i gpurun_out/r01i_accum.ncu-rep --page raw --csv > gpurun_out/r01i_accum_raw.csv 2>/dev/null
  ncu -i gpurun_out/r01i_accum.ncu-rep --page source --csv --print-source sass > gpurun_out/r01i_accum_source.csv 2>/dev/null
  python tools/ncu_source_top.py gpurun_out/r01i_accum_source.csv --top 40 > gpurun_out/r01i_accum_source_top.txt 2>&1
  gzip -9 -f gpurun_out/r01i_accum_source.csv
  rm -f gpurun_out/r01i_accum.ncu-rep
fi
rm -f gpurun_out/r01h_ntt_source.csv gpurun_out/r01h_ntt.ncu-rep
du -sm gpurun_out; NOW=$(date +%s); echo "elapsed $((NOW-T0)) s"
