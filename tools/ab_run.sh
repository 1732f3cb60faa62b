#!/bin/bash
# A/B harness: tools/ab_run.sh "<lib or ->:<EMC>:<EMV>[:extra env]" ...   -> gpurun_out/ab.jsonl
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r lib emc emv extra <<<"$spec"
  envs="MC33_B200_EMC_PER_SM=$emc MC33_B200_EMV_PER_SM=$emv $extra"
  if [ "$lib" != "-" ]; then envs="$envs MC33_B200_LIB=$PWD/$lib"; fi
  env $envs python tools/time_pipeline.py 512 3 "$spec" >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
done
cat gpurun_out/ab.jsonl
