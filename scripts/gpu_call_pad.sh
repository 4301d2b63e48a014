#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pad.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_pad.txt
ADER_B200_TRACE=gpurun_out/trace_pad.json timeout 120 python bench.py > gpurun_out/bench_pad.json 2> gpurun_out/bench_pad.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_pad.txt
ADER_B200_DEFINES=-DADER_TC_TIMELINE python -c "from ader_b200 import build; build.build(force=True)" > gpurun_out/tl_build.log 2>&1
ADER_B200_STEP_IMPL=groups timeout 120 python scripts/tc_timeline.py > gpurun_out/tc_timeline_pad.txt 2>&1
tail -4 gpurun_out/pytest_pad.log; cat gpurun_out/legs_pad.txt
python - <<PY
import json
d=json.load(open("gpurun_out/bench_pad.json")); print(round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["tc_kernels_ms"], d["loss_group_ms"], d["roofline"]["frac"])
PY
grep "^==" gpurun_out/tc_timeline_pad.txt
