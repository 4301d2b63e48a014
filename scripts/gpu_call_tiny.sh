#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 24 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resume or herding_periods" > gpurun_out/pytest_tiny.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/pytest_tiny.log
