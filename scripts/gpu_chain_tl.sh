#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
ADER_B200_DEFINES=-DADER_TC_TIMELINE python -c "from ader_b200 import build; build.build(force=True)" && python scripts/fz_timeline.py > gpurun_out/chain_timeline.txt 2>&1
cat gpurun_out/chain_timeline.txt | tail -40
