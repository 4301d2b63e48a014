#!/bin/bash
# bench.py at N GPUs the way the driver launches it (torchrun, one rank per GPU); both gradient back ends.
N=${1:-8}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
: > gpurun_out/legs_scale_$N.txt
leg() { echo "$1 rc=$2 t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_scale_$N.txt; }
for be in ${2:-auto nccl}; do
  ADER_B200_DP=$be timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n${N}_$be.json 2> gpurun_out/bench_n${N}_$be.err; leg bench_n${N}_$be $?
done
cat gpurun_out/legs_scale_$N.txt
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n${N}_*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 4),
              "dp", d.get("dp_backend"), "strong", (d.get("strong_scaling") or {}).get("ms_per_step"),
              "k_dp_adam us", d["kernels_us_per_step"].get("ader::dp::k_dp_adam"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
grep -h "WATCHDOG\|Error\|error" gpurun_out/bench_n${N}_*.err | head -10
