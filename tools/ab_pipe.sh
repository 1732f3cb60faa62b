#!/bin/bash
# tools/ab_pipe.sh "<env settings>" ... -- workloads...   e.g.  tools/ab_pipe.sh "MC33_B200_PIPE=1" "MC33_B200_PIPE=2" -- cfg2 cfg3
# one tools/time_workload.py line per (env set, workload) -> gpurun_out/ab_pipe.txt
mkdir -p gpurun_out
envs=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
shift
for w in "$@"; do
  for e in "${envs[@]}"; do
    tag=$(echo "$e" | sed 's#[^A-Za-z0-9]#_#g' | tail -c 60)
    env $e timeout 600 python tools/time_workload.py $w > gpurun_out/tw_${w}_${tag}.json 2> gpurun_out/tw_${w}_${tag}.err
    echo "$w [$e] $(head -c 700 gpurun_out/tw_${w}_${tag}.json)" >> gpurun_out/ab_pipe.txt
  done
done
cat gpurun_out/ab_pipe.txt
