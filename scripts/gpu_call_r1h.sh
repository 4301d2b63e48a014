#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_h.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_h.txt
ADER_B200_TRACE=gpurun_out/trace_h.json timeout 120 python bench.py > gpurun_out/bench_h_default.json 2> gpurun_out/bench_h_default.err
echo "bench default rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_h.txt
ADER_B200_DAG_PRIO=0 ADER_B200_TRACE=gpurun_out/trace_h_noprio.json timeout 120 python bench.py > gpurun_out/bench_h_noprio.json 2> gpurun_out/bench_h_noprio.err
echo "bench noprio rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_h.txt
tail -6 gpurun_out/pytest_h.log
cat gpurun_out/legs_h.txt
for f in default noprio; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_h_$f.json")); print("$f", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["e2e"]["last_loss"], d["gpu_launches_per_step"])
except Exception as e: print("$f failed", e)
PY
done
