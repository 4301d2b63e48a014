#!/bin/bash
# scatter second generation (shared-memory rows, rolled loop), batched position-gradient / partial-reduction loads: parity + A/B
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_encoder_fused.py tests/test_gpu_dp.py tests/test_gpu_parity.py -x -q -k "not end_to_end and not resume and not herding and not fisher and not eval" > gpurun_out/r2b_pytest_quick.log 2>&1
echo "pytest quick rc=$? t=$(( $(date +%s) - T0 ))" > gpurun_out/r2b_legs.txt
tail -3 gpurun_out/r2b_pytest_quick.log
run() {   # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --no-period > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_$name.json"))
    print("$name", round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["gpu_launches_per_step"], {k.replace("ader::", ""): v for k, v in list(d["kernels_us_per_step"].items())[:5]})
except Exception as e:
    print("$name failed", e)
PY
}
run default ADER_B200_TRACE=gpurun_out/r2b_trace.json
python scripts/trace_summary.py gpurun_out/r2b_trace.json > gpurun_out/r2b_timeline.txt 2>&1; rm -f gpurun_out/r2b_trace.json
run scatter1 ADER_B200_SCATTER=1
run nofuse ADER_B200_FUSE_DREP=0
run wgrad1 ADER_B200_WGRAD=1
run default2
cat gpurun_out/r2b_legs.txt
tail -28 gpurun_out/r2b_timeline.txt
