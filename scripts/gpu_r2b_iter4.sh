#!/bin/bash
# tail reorder (scatter first, grid-stride untouched-row Adam) + config-5 full step at N = 1
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_encoder_fused.py tests/test_gpu_dp.py -x -q > gpurun_out/r2b_pytest_quick.log 2>&1
echo "pytest quick rc=$? t=$(( $(date +%s) - T0 ))" > gpurun_out/r2b_legs.txt
tail -3 gpurun_out/r2b_pytest_quick.log
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
run nosplit ADER_B200_SPLIT_ADAM=0
run default2
timeout 300 python bench.py --config synthetic1m --steps 20 --warmup 3 > gpurun_out/r2b_bench_synth1m_n1.json 2> gpurun_out/r2b_bench_synth1m_n1.err
echo "synth1m rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/r2b_legs.txt
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_synth1m_n1.json"))
    print("synth1m", round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3), d["value"], d["roofline"]["frac"], d["phases_ms_eager"], d["kernels_us_per_step"])
except Exception as e:
    print("synth1m failed", e)
PY
tail -5 gpurun_out/r2b_bench_synth1m_n1.err
cat gpurun_out/r2b_legs.txt
tail -22 gpurun_out/r2b_timeline.txt
