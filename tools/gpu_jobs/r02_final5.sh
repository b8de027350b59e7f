#!/bin/bash
# Round 2, last check of the committed defaults (lazily reduced additions + dedicated squaring): the default bench line,
# the whole GPU suite, smoke.
set -u
mkdir -p gpurun_out
echo "== bench (default flags)"
SECONDS=0
timeout 300 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$? wall=${SECONDS}s"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "prove_ms_serial", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["one_in_flight"], d["e2e"]["two_in_flight"])
print(d["roofline"]["frac"], d["roofline"]["practical_peak_frac"], d["roofline"]["executed_frac_of_plain_imad_peak"], d["roofline"]["avg_launch_ms"], d["roofline_ntt"]["h_pipeline"]["ms"])
print(d["gpu_witness"]["solve_ms"], d["batch_2p22"]["proofs_per_s"], d["cpu_baseline"]["seconds_per_proof"], d["cpu_baseline"]["matches_gpu_proof"], d["clocks"])
PY
echo "== pytest -m gpu"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
