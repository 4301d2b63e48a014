#!/bin/bash
# last check of the round on the final library: full GPU suite, smoke(), one short bench line
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_last.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2b_pytest_last.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 python bench.py --no-period --steps 100 > gpurun_out/r2b_bench_last.json 2> gpurun_out/r2b_bench_last.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench_last.json"))
print(round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["roofline_step_dominant"])
PY
