#!/bin/bash
# Round 2, job q: the small-circuit bench lines (BASELINE configs[0]) of the final code (lazily reduced additions).
set -u
mkdir -p gpurun_out
for shape in tx withdraw; do
  timeout 300 python bench.py --shape $shape --no-batch-2p22 --steps 30 > gpurun_out/r02_bench_${shape}_n1.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_${shape}_n1.json').read().strip().splitlines()[-1]); e=d['e2e']
print('$shape', d['ms_per_step'], e['value'], e['one_in_flight'], e['two_in_flight'], d['prove_ms_serial'], d['gpu_witness']['solve_ms'], d['cpu_baseline']['seconds_per_proof'], d['cpu_baseline']['matches_gpu_proof'])"
done
