#!/bin/bash
# Third pass: parity with the windowed scatter / priority chain, then A/B bench lines of the experiment switches.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_f.txt
ADER_B200_PDL=1 timeout 200 python -m pytest tests/test_gpu_encoder_fused.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_f_pdl.log 2>&1
echo "pytest pdl rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_f.txt
ADER_B200_TRACE=gpurun_out/trace_default.json timeout 120 python bench.py > gpurun_out/bench_f_default.json 2> gpurun_out/bench_f_default.err
echo "bench default rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_f.txt
ADER_B200_PDL=1 ADER_B200_TRACE=gpurun_out/trace_pdl.json timeout 120 python bench.py > gpurun_out/bench_f_pdl.json 2> gpurun_out/bench_f_pdl.err
echo "bench pdl rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_f.txt
ADER_B200_DAG_PRIO=0 timeout 120 python bench.py > gpurun_out/bench_f_noprio.json 2> gpurun_out/bench_f_noprio.err
echo "bench noprio rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_f.txt
ADER_B200_SCATTER=legacy timeout 120 python bench.py > gpurun_out/bench_f_legacy.json 2> gpurun_out/bench_f_legacy.err
echo "bench legacy-scatter rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_f.txt
tail -4 gpurun_out/pytest_f.log; tail -4 gpurun_out/pytest_f_pdl.log
cat gpurun_out/legs_f.txt
for f in default pdl noprio legacy; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_f_$f.json")); print("$f", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["e2e"]["last_loss"])
except Exception as e: print("$f failed", e)
PY
done
