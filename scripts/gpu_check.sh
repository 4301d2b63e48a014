#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_check.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_check.txt
timeout 120 python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_check.txt
tail -15 gpurun_out/pytest_check.log; cat gpurun_out/legs_check.txt
python - <<PY
import json
d=json.load(open("gpurun_out/bench_check.json")); print(round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["roofline"]["frac"])
PY
