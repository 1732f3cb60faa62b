#!/bin/bash
# tools/ab_env.sh "VAR=val VAR2=val" ...  -> one time_pipeline line per env set (gpurun_out/ab.jsonl)
mkdir -p gpurun_out
for spec in "$@"; do
  env $spec python tools/time_pipeline.py 512 3 "$spec" >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
done
python - <<'PY'
import json
for l in open("gpurun_out/ab.jsonl"):
    d = json.loads(l); print(d["tag"], d["ms_per_iso"], d["serial_kernel_ms"])
PY
