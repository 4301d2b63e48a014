#!/bin/bash
# one optimisation iteration on the tc2 kernels: parity tests, bench line, then the per-CTA timeline (instrumented rebuild)
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc_loss_path or tc_full_size or vocab_parallel or default_tc_path" > gpurun_out/pytest_tc_iter.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tc_iter.log
timeout 150 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_tc_iter.json 2> gpurun_out/bench_tc_iter.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/bench_tc_iter.json") if l.startswith("{")][-1]
    print("ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "tc_kernels_ms", round(d["tc_kernels_ms"], 4),
          "loss_group_ms", round(d["loss_group_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
    print({k: v for k, v in d["kernels_us_per_step"].items()})
except Exception as ex:
    print("bench unreadable", ex)
PY
ADER_B200_DEFINES=-DADER_TC_TIMELINE python -c "from ader_b200 import build; build.build(force=True)" && python scripts/tc2_timeline.py > gpurun_out/tc2_timeline.txt 2>&1
grep -A8 "==\|three" gpurun_out/tc2_timeline.txt | head -60
