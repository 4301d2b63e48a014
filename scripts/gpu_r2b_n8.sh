#!/bin/bash
# one 8-GPU box: config 5 at N = 4 and 8 (DP), then the headline config at N = 8 with the round's final kernels
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
bash scripts/gpu_r2b_synth_dp.sh 4 2>&1 | tail -3
bash scripts/gpu_r2b_synth_dp.sh 8 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 8 > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err
echo "default N=8 rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2b_bench_n8.json") if l.startswith("{")][-1])
    print("default N=8", round(d["ms_per_step"], 4), d["value"], d["dp_backend"], d["strong_scaling"]["ms_per_step"])
except Exception as e:
    print("failed", e)
PY
