#!/bin/bash
# single-CTA packing kernel: parity + A/B against the three-kernel form
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_encoder_fused.py tests/test_gpu_dp.py tests/test_gpu_parity.py -x -q -k "not end_to_end and not resume and not herding and not fisher and not eval" > gpurun_out/r2b_pytest_quick.log 2>&1
echo "pytest quick rc=$?"; tail -2 gpurun_out/r2b_pytest_quick.log
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --no-period > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_$name.json"))
    print("$name", round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["gpu_launches_per_step"])
except Exception as e:
    print("$name failed", e)
PY
}
run pack1a
run pack3a ADER_B200_PACK=3
run pack1b
run pack3b ADER_B200_PACK=3
