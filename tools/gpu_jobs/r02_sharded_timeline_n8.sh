#!/bin/bash
N=8 bash tools/gpu_jobs/r02_sharded_timeline.sh
