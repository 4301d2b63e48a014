#!/bin/bash
# config 5 (1M items) as a full data-parallel step at N GPUs: weak line + strong-scaling object
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --config synthetic1m --steps 20 --warmup 3 > gpurun_out/r2b_bench_synth1m_n$N.json 2> gpurun_out/r2b_bench_synth1m_n$N.err
echo "rc=$?"
grep -v "^$" gpurun_out/r2b_bench_synth1m_n$N.err | tail -12
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2b_bench_synth1m_n$N.json") if l.startswith("{")][-1])
    print("synth1m N=$N", round(d["ms_per_step"], 3), d["value"], d["dp_backend"], d["strong_scaling"])
except Exception as e:
    print("failed", e)
PY
