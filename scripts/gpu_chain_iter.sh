#!/bin/bash
# iteration leg for the chained encoder kernels: parity tests, bench with / without the chain, CUPTI step timelines
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_encoder_fused.py -m gpu -x -q > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_chain.log
for c in 1 0; do
  ADER_B200_CHAIN=$c timeout 300 python bench.py --no-period > gpurun_out/bench_chain_$c.json 2> gpurun_out/bench_chain_$c.err; echo "bench chain=$c rc=$?"
  python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_chain_$c.json") if l.startswith("{")][-1]
print("chain=$c ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "launches", d.get("gpu_launches"))
PY
  ADER_B200_CHAIN=$c ADER_B200_TRACE=gpurun_out/trace_chain.json timeout 300 python bench.py --no-period --steps 50 --warmup 5 > /dev/null 2>&1
  python scripts/trace_summary.py gpurun_out/trace_chain.json > gpurun_out/timeline_chain_$c.txt 2>&1; rm -f gpurun_out/trace_chain.json
  cat gpurun_out/timeline_chain_$c.txt
done
