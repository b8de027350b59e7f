#!/bin/bash
N=2 bash tools/gpu_jobs/r02_sharded_timeline.sh
