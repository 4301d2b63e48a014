#!/bin/bash
# parity + A/B of: k_wgrad2, split table update (untouched rows beside the scatter), fused d_rep reduction
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" > gpurun_out/r2b_legs.txt
tail -5 gpurun_out/r2b_pytest.log
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --no-period > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_$name.json"))
    print("$name", round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["gpu_launches_per_step"], {k.replace("ader::", ""): v for k, v in list(d["kernels_us_per_step"].items())[:6]})
except Exception as e:
    print("$name failed", e)
PY
}
run default ADER_B200_TRACE=gpurun_out/r2b_trace.json
python scripts/trace_summary.py gpurun_out/r2b_trace.json > gpurun_out/r2b_timeline.txt 2>&1; rm -f gpurun_out/r2b_trace.json
run wgrad1 ADER_B200_WGRAD=1
run nosplit ADER_B200_SPLIT_ADAM=0
run nofuse ADER_B200_FUSE_DREP=0
run old ADER_B200_WGRAD=1 ADER_B200_SPLIT_ADAM=0 ADER_B200_FUSE_DREP=0
cat gpurun_out/r2b_legs.txt
head -60 gpurun_out/r2b_timeline.txt
