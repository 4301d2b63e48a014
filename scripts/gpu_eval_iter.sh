#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_parity.py tests/test_protocol_product.py -m gpu -x -q -k "eval or evaluator or end_to_end or resume or epoch_queue" > gpurun_out/pytest_evaliter.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_evaliter.log
timeout 400 python bench.py > gpurun_out/bench_evaliter.json 2> gpurun_out/bench_evaliter.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_evaliter.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_evaliter.json") if l.startswith("{")][-1]
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
for p in d["period_metric"]["periods"]:
    print("   period", p["period"], "steady", round(p["steady_sessions_per_s"]), "eval rows/s", round(p["eval_rows_per_s"]), "eval_s", p["eval_s"], "rows", p["eval_rows"])
PY
