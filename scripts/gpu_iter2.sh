#!/bin/bash
# iteration leg: eval tests, tc2 parity tests + bench + timeline
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_eval.py -m gpu -x -q > gpurun_out/pytest_eval.log 2>&1; echo "eval pytest rc=$?"; tail -25 gpurun_out/pytest_eval.log
bash scripts/gpu_tc2_iter.sh
