#!/bin/bash
# ncu --set full captures (one launch per kernel, after warm-up) of a bench workload.  Usage: gpu_profile.sh TAG WORKLOAD "kernel regex ..." 
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=$1; WL=$2; shift 2
for KR in "$@"; do
  N=$(echo "$KR" | tr -c 'a-zA-Z0-9' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KR" -s 6 -c 3 -f -o gpurun_out/prof_${TAG}_${WL}_$N \
      python bench.py --workload $WL --steps 6 --warmup 3 --resident-frames 10 --cpu-budget 0 > gpurun_out/ncu_${TAG}_${WL}_$N.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_${WL}_$N.log | cut -c1-200
done
ls -la gpurun_out/prof_${TAG}_* 
