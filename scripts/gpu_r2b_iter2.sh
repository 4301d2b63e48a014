#!/bin/bash
# k_wgrad phase stamps (timeline build) + A/B of the two tail switches
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
ADER_B200_LIB=ader_b200/lib/libader_b200_tl.so timeout 120 python scripts/wgrad_timeline.py > gpurun_out/r2b_wgrad_tl.txt 2>&1
for v in "0 1" "1 0" "0 0" "1 1"; do
  set -- $v
  ADER_B200_SPLIT_ADAM=$1 ADER_B200_FUSE_DREP=$2 timeout 200 python bench.py --no-period > gpurun_out/r2b_bench_s$1_f$2.json 2> gpurun_out/r2b_bench_s$1_f$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_s$1_f$2.json"))
    print("split=$1 fuse=$2", round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["gpu_launches_per_step"])
except Exception as e:
    print("split=$1 fuse=$2 failed", e)
PY
done
cat gpurun_out/r2b_wgrad_tl.txt
