#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
ADER_B200_DEFINES=-DADER_TC_TIMELINE python -c "from ader_b200 import build; build.build(force=True)" > gpurun_out/tl_build.log 2>&1
ADER_B200_STEP_IMPL=groups timeout 120 python scripts/tc_timeline.py > gpurun_out/tc_timeline.txt 2>&1
cat gpurun_out/tc_timeline.txt | cut -c1-900
