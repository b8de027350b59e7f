#!/bin/bash
# Round-2 opener (prepared at the end of round 1, which ran out of GPU minutes before this capture):
# source-level ncu view of the two accumulation kernels (54 % of a proof's kernel time) and of the boundary levels,
# plus NTT passes of the final round-1 code.  ~4 GPU-minutes.  Summaries: tools/ncu_summary.py raw ..., tools/ncu_source_top.py.
set -u
mkdir -p gpurun_out
# -k takes the function base name (no template arguments): capture the 5 level-1 launches of the second proof
echo "== ncu full: k_accum_affine (4 x G1 + 1 x G2 of one proof)"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_accum_affine -s 5 -c 5 -f -o gpurun_out/r02a_accum \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02a_ncu_accum.log 2>&1; tail -c 150 gpurun_out/r02a_ncu_accum.log
ncu -i gpurun_out/r02a_accum.ncu-rep --page raw --csv > gpurun_out/r02a_accum_raw.csv 2>/dev/null
ncu -i gpurun_out/r02a_accum.ncu-rep --page source --csv --print-source sass > gpurun_out/r02a_accum_source.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/r02a_accum_source.csv --top 40 > gpurun_out/r02a_accum_source_top.txt 2>&1
gzip -9 -f gpurun_out/r02a_accum_source.csv; rm -f gpurun_out/r02a_accum.ncu-rep
echo "== ncu full: NTT passes of the final round-1 code (2^20)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 8 -f -o gpurun_out/r02a_ntt \
    python tools/ntt_probe.py --log-n 20 --reps 1 --no-time > gpurun_out/r02a_ncu_ntt.log 2>&1; tail -c 150 gpurun_out/r02a_ncu_ntt.log
ncu -i gpurun_out/r02a_ntt.ncu-rep --page raw --csv > gpurun_out/r02a_ntt_raw.csv 2>/dev/null
ncu -i gpurun_out/r02a_ntt.ncu-rep --page source --csv --print-source sass > gpurun_out/r02a_ntt_source.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/r02a_ntt_source.csv --top 25 > gpurun_out/r02a_ntt_source_top.txt 2>&1
gzip -9 -f gpurun_out/r02a_ntt_source.csv; rm -f gpurun_out/r02a_ntt.ncu-rep
echo "== experiment: G2 accumulation with ZZ / ZZZ in shared memory (168 registers, 3 CTAs/SM): parity, then A/B"
ZKR_RUN_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k smem_accumulator 2>&1 | tail -3
timeout 200 python bench.py --no-cpu --steps 10 > gpurun_out/r02a_bench_default.json 2>/dev/null
ZKR_G2_SMEM_ACC=1 timeout 200 python bench.py --no-cpu --steps 10 > gpurun_out/r02a_bench_g2smz.json 2>/dev/null
python - <<PY
import json
for f in ("gpurun_out/r02a_bench_default.json", "gpurun_out/r02a_bench_g2smz.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["prove_ms_serial"], d["stage_ms_overlapped"]["msm_b2_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
du -sm gpurun_out
