#!/bin/bash
# usage (under gpurun): bash profiles/run_profile.sh <tag>   -> gpurun_out/prof_<tag>.ncu-rep + launches_<tag>.csv
TAG=${1:-x}
WL=${2:-c5}
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --chunk 1 --e2e-steps 2"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:uaisa_env -s 6 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log
