#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_encoder_fused.py -m gpu -x -q -k "epoch_queue or graph_step or end_to_end_three" > gpurun_out/pytest_group.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_group.log
for g in 4 1 8; do
  ADER_B200_GRAPH_STEPS=$g timeout 400 python bench.py > gpurun_out/bench_group_$g.json 2> gpurun_out/bench_group_$g.err; echo "bench group=$g rc=$?"; tail -2 gpurun_out/bench_group_$g.err
  python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_group_$g.json") if l.startswith("{")][-1]
print("group=$g ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
for p in d["period_metric"]["periods"]:
    print("   period", p["period"], "steady", round(p["steady_sessions_per_s"]), "epoch_s", p["epoch_s"], "launch_s", p["host_launch_s"])
PY
done
