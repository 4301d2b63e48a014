#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_last2.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_last2.txt
R=gpurun_out/results_last2; mkdir -p $R
( time timeout 150 python -m ader_b200.main --data_root=data_cache --results_root $R --dataset=DIGINETICA --save_dir=ADER ) > $R/diginetica_ader.log 2>&1
echo "diginetica rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_last2.txt
tail -5 gpurun_out/pytest_last2.log; cat gpurun_out/legs_last2.txt
grep -E "train throughput|Average|Total time|real|Error|error" $R/diginetica_ader.log | cut -c1-330
