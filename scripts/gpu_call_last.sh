#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_last.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_last.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1
echo "smoke rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_last.txt
timeout 120 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_last.txt
R=gpurun_out/results_last; mkdir -p $R
( time timeout 200 python -m ader_b200.main --data_root=data_cache --cache_dir=gpurun_out/pair_cache --results_root $R --dataset=DIGINETICA --save_dir=ADER ) > $R/diginetica_ader.log 2>&1
echo "diginetica rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_last.txt
tail -5 gpurun_out/pytest_last.log; tail -2 gpurun_out/smoke_last.log; cat gpurun_out/legs_last.txt
python - <<PY
import json
d=json.load(open("gpurun_out/bench_last.json")); print(round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["roofline"]["frac"])
PY
grep -E "train throughput|Average|Total time|real" $R/diginetica_ader.log | cut -c1-260
