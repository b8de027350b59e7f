#!/bin/bash
# Round 2, job g: witness solver tests, launch list of a standalone 2^24 G1 MSM with the gather forced (why is it slow
# at 2M buckets?), full sweep with the G2 infinity check fixed.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest witness + msm + ntt"
timeout 900 python -m pytest tests/test_gpu_witness.py tests/test_gpu_msm.py tests/test_gpu_ntt.py -m gpu -x -q --durations=5 2>&1 | tail -12
health after-tests
echo "== ncu launch list: G1 MSM 2^24, gather forced"
ZKR_MSM_LEVELS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02g_msm24_gather.csv \
    python tools/sweep.py --min-log 24 --max-log 24 --g2-max-log 0 --skip-ntt --reps 1 --out gpurun_out/r02g_tmp.json > /dev/null 2>&1
python - <<PY
import csv, collections, re
rows = list(csv.reader(open("gpurun_out/r02g_msm24_gather.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) > vi and r[vi]:
        k = re.sub(r"\(.*", "", r[ki]).replace("void zkr::", "").replace("void ", "")[:60]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%10.1f us %3d %s" % (v / 1000, c, k))
PY
health after-ncu
echo "== sweep with CPU baselines (final)"
timeout 1500 python bench.py --sweep msm,ntt --sweep-max-log 26 --sweep-out gpurun_out/r02g_sweep_1gpu.json > gpurun_out/r02g_sweep.log 2> gpurun_out/r02g_sweep.err; echo sweep rc=$?; tail -3 gpurun_out/r02g_sweep.err; grep -c '"correct": true' gpurun_out/r02g_sweep.log; grep -c '"correct": false' gpurun_out/r02g_sweep.log
health end
