#!/bin/bash
# Round 2, check of the code with the lazily reduced additions as the default: whole GPU suite, smoke, the default bench
# line, the launch list and a --set full capture of the accumulation kernels.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
health after-tests
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags)"
SECONDS=0
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$? wall=${SECONDS}s"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "prove_ms_serial", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["one_in_flight"], d["e2e"]["two_in_flight"])
print(d["roofline"]["frac"], d["roofline"]["practical_peak_frac"], d["roofline_ntt"]["transform"]["ms"], d["roofline_ntt"]["h_pipeline"]["ms"])
print(d["gpu_witness"]["solve_ms"], d["batch_2p22"]["proofs_per_s"], d["cpu_baseline"]["seconds_per_proof"], d["cpu_baseline"]["matches_gpu_proof"], d["clocks"])
PY
echo "== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-batch-2p22 --no-gpu-witness --no-two-in-flight > gpurun_out/r02_ncu_launch_bench.log 2>&1; tail -c 100 gpurun_out/r02_ncu_launch_bench.log
echo "== ncu full: accumulation kernels (G1 x4 + G2) of the second proof"
timeout 300 ncu --set full --clock-control none -k regex:"k_accum_affine" -s 5 -c 5 -f -o gpurun_out/r02_msm \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-batch-2p22 --no-gpu-witness --no-two-in-flight > gpurun_out/r02_ncu_msm.log 2>&1; tail -c 100 gpurun_out/r02_ncu_msm.log
ncu -i gpurun_out/r02_msm.ncu-rep --page raw --csv > gpurun_out/r02_msm_lazy_raw.csv 2>/dev/null
rm -f gpurun_out/r02_msm.ncu-rep
python tools/ncu_summary.py raw gpurun_out/r02_msm_lazy_raw.csv gpurun_out/r02_ncu_accum_lazy_summary.md --traffic-json gpurun_out/r02_ncu_traffic_lazy.json > /dev/null 2>&1; ls -la gpurun_out/r02_ncu_accum_lazy_summary.md gpurun_out/r02_ncu_traffic_lazy.json
health end
