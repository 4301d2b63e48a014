#!/bin/bash
# tc2 (TMA + SWIZZLE_128B) loss kernels: parity tests per generation / gradient-product width, then the bench line of each.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
: > gpurun_out/legs_tc2.txt
T0=$(date +%s)
leg() { echo "$1 rc=$2 t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_tc2.txt; }
for cfg in "2 160" "2 192"; do
  set -- $cfg
  ADER_B200_TC=$1 ADER_B200_TC2_N2=$2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "tc_loss_path or tc_full_size or vocab_parallel or default_tc_path" > gpurun_out/pytest_tc_$1_$2.log 2>&1; leg "pytest_tc gen=$1 n2=$2" $?
  ADER_B200_TC=$1 ADER_B200_TC2_N2=$2 timeout 150 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_tc_$1_$2.json 2> gpurun_out/bench_tc_$1_$2.err; leg "bench gen=$1 n2=$2" $?
done
cat gpurun_out/legs_tc2.txt
for f in gpurun_out/pytest_tc_*.log; do echo "== $f"; tail -5 $f; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_tc_*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "tc_kernels_ms", round(d["tc_kernels_ms"], 4),
              "loss_group_ms", round(d["loss_group_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
        print("    ", {k: v for k, v in d["kernels_us_per_step"].items() if "tc" in k})
    except Exception as ex:
        print(f, "unreadable", ex)
PY
