#!/bin/bash
# Multi-GPU check (gpurun --gpus N): GPU test suite (incl. the 2-GPU data-parallel parity tests), then bench.py at
# 1 GPU and at N GPUs with both gradient back ends.  Every leg is bounded by its own timeout.
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
leg() { echo "$1 rc=$2 t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_dp.txt; }
: > gpurun_out/legs_dp.txt
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_dp.log 2>&1; leg pytest $?
timeout 150 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; leg bench_n1 $?
for be in p2p nccl; do
  ADER_B200_DP=$be timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n${N}_$be.json 2> gpurun_out/bench_n${N}_$be.err; leg bench_n${N}_$be $?
done
tail -12 gpurun_out/pytest_dp.log; cat gpurun_out/legs_dp.txt
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n*.json")):
    try:
        d = json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 4),
              "frac", round(d["roofline"]["frac"], 4), "dp", d.get("dp_backend"), "strong", (d.get("strong_scaling") or {}).get("ms_per_step"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
grep -h "bench rank" gpurun_out/bench_n${N}_*.err | tail -30
