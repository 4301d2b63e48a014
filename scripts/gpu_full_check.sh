#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_full.log
timeout 400 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_full.json") if l.startswith("{")][-1]
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "tc_kernels_ms", d["tc_kernels_ms"])
print("period", json.dumps(d["period_metric"], indent=0)[:1500])
print("components", d["components"])
print("cpu", d["cpu_components"])
PY
