#!/bin/bash
# Fourth pass: parity (fused step with Adam in the DAG, PDL default on), bench + timeline.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_g.txt
ADER_B200_TRACE=gpurun_out/trace_g.json timeout 120 python bench.py > gpurun_out/bench_g_default.json 2> gpurun_out/bench_g_default.err
echo "bench default rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_g.txt
ADER_B200_PDL=0 timeout 120 python bench.py > gpurun_out/bench_g_nopdl.json 2> gpurun_out/bench_g_nopdl.err
echo "bench nopdl rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_g.txt
tail -6 gpurun_out/pytest_g.log
cat gpurun_out/legs_g.txt
for f in default nopdl; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_g_$f.json")); print("$f", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["e2e"]["last_loss"], d["gpu_launches_per_step"])
except Exception as e: print("$f failed", e)
PY
done
tail -3 gpurun_out/bench_g_default.err
