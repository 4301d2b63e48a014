#!/bin/bash
# round-2 second session: parity of the new tail (split Adam, fused d_rep reduction, deeper weight-gradient ring) + A/B bench
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" > gpurun_out/r2b_legs.txt
ADER_B200_TRACE=gpurun_out/r2b_trace.json timeout 200 python bench.py --no-period > gpurun_out/r2b_bench_new.json 2> gpurun_out/r2b_bench_new.err
echo "bench new rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/r2b_legs.txt
python scripts/trace_summary.py gpurun_out/r2b_trace.json > gpurun_out/r2b_timeline.txt 2>&1; rm -f gpurun_out/r2b_trace.json
ADER_B200_SPLIT_ADAM=0 ADER_B200_FUSE_DREP=0 timeout 200 python bench.py --no-period > gpurun_out/r2b_bench_oldtail.json 2> gpurun_out/r2b_bench_oldtail.err
echo "bench oldtail rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/r2b_legs.txt
tail -5 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_legs.txt
python - <<PY
import json
for n in ("new", "oldtail"):
    try:
        d = json.load(open("gpurun_out/r2b_bench_%s.json" % n))
        print(n, round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), round(d["roofline"]["frac"], 4), d["gpu_launches_per_step"], d["kernels_us_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
head -60 gpurun_out/r2b_timeline.txt
