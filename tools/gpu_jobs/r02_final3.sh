#!/bin/bash
# Round 2, final check of the committed code: whole GPU suite, smoke, the small-circuit bench lines, the default bench line.
set -u
mkdir -p gpurun_out
health() { timeout 30 nvidia-smi --query-gpu=name,memory.used,utilization.gpu --format=csv,noheader; echo "health rc=$? ($1)"; }
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
health after-tests
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for shape in tx withdraw; do
  timeout 300 python bench.py --shape $shape --no-batch-2p22 --steps 30 > gpurun_out/r02_bench_${shape}_n1.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_${shape}_n1.json').read().strip().splitlines()[-1]); e=d['e2e']
print('$shape', d['ms_per_step'], e['value'], e['one_in_flight'], e['two_in_flight'], d['prove_ms_serial'], d['gpu_witness']['solve_ms'], d['cpu_baseline']['seconds_per_proof'])"
done
echo "== bench (default flags)"
SECONDS=0
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$? wall=${SECONDS}s"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "prove_ms_serial", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["one_in_flight"], d["e2e"]["two_in_flight"])
print(d["roofline"]["frac"], d["roofline_ntt"]["transform"]["ms"], d["roofline_ntt"]["h_pipeline"], d["roofline_ntt"]["avg_launch_ms"])
print(d["gpu_witness"]["solve_ms"], d["batch_2p22"]["proofs_per_s"], d["cpu_baseline"]["seconds_per_proof"], d["clocks"])
PY
health end
