#!/bin/bash
# usage: tools/gpu_jobs/run.sh <job-name> [gpurun args...]   -- runs tools/gpu_jobs/<job>.sh on a GPU box, retrying while the pod is busy
job=$1; shift
log=gpurun_out/${job}_call.log
for try in 1 2 3 4 5 6 7 8; do
    /usr/local/graft/bin/gpurun --timeout 2700 "$@" -- "bash tools/gpu_jobs/$job.sh" > $log 2>&1
    if grep -q "status=transient\|status=busy\|retry in a few minutes" $log; then sleep 150; continue; fi
    break
done
echo done >> $log
