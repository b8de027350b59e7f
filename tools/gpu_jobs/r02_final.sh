#!/bin/bash
# Round 2, final single-GPU job: the whole GPU suite, the default bench line (the record the driver will reproduce),
# launch list and --set full captures of the final kernels.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14
health after-tests
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags)"
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo rc=$?
grep "\[bench\]\|Elapsed" gpurun_out/r02_bench_n1.err | tail -8
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "prove_ms_serial", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["one_in_flight"], d["e2e"]["two_in_flight"])
print(d["roofline"]["frac"], d["roofline"]["practical_peak_frac"], d["roofline_ntt"]["transform"], d["roofline_ntt"]["h_pipeline"])
print(d["gpu_witness"]); print(d["batch_2p22"]["proofs_per_s"], d["cpu_baseline"]["seconds_per_proof"], d["clocks"])
PY
echo "== reference arm (as the driver runs it)"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -1 gpurun_out/r02_bench_reference.json | cut -c1-400
echo "== witness solve: persistent vs per-level"
ZKR_WITNESS_PER_LEVEL=1 timeout 300 python bench.py --no-cpu --no-batch-2p22 --no-two-in-flight --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('per-level', d['gpu_witness']['solve_ms'], d['gpu_witness']['solve_plus_prove_ms'])"
health after-bench
echo "== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-batch-2p22 --no-gpu-witness --no-two-in-flight > gpurun_out/r02_ncu_launch_bench.log 2>&1; tail -c 100 gpurun_out/r02_ncu_launch_bench.log
echo "== ncu full: accumulation (G1 x4 + G2) and gather / reduction kernels of the second proof"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_accum_affine|k_bucket_gather|k_bucket_sums|k_bucket_weighted" -s 20 -c 20 -f -o gpurun_out/r02_msm \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-batch-2p22 --no-gpu-witness --no-two-in-flight > gpurun_out/r02_ncu_msm.log 2>&1; tail -c 100 gpurun_out/r02_ncu_msm.log
ncu -i gpurun_out/r02_msm.ncu-rep --page raw --csv > gpurun_out/r02_msm_raw.csv 2>/dev/null
rm -f gpurun_out/r02_msm.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:k_ntt_pass -c 6 -f -o gpurun_out/r02_ntt \
    python tools/ntt_probe.py --log-n 20 --reps 1 --no-time > gpurun_out/r02_ncu_ntt.log 2>&1
ncu -i gpurun_out/r02_ntt.ncu-rep --page raw --csv > gpurun_out/r02_ntt_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ntt.ncu-rep
python tools/ncu_summary.py raw gpurun_out/r02_msm_raw.csv gpurun_out/r02_ntt_raw.csv gpurun_out/r02_ncu_full_summary.md --traffic-json gpurun_out/r02_ncu_traffic.json > /dev/null 2>&1; ls -la gpurun_out/r02_ncu_full_summary.md gpurun_out/r02_ncu_traffic.json
health end
du -sm gpurun_out
