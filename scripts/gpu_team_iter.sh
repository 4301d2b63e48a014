#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_encoder_fused.py tests/test_gpu_parity.py -m gpu -x -q -k "fused or chained or default_tc or train_grad or dropout" > gpurun_out/pytest_team.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_team.log
for c in 1 0; do
  ADER_B200_TEAM_ATTN=$c timeout 300 python bench.py --no-period > gpurun_out/bench_team_$c.json 2> gpurun_out/bench_team_$c.err; echo "bench team=$c rc=$?"
  python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/bench_team_$c.json") if l.startswith("{")][-1]
print("team=$c ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "launches", d.get("gpu_launches"))
PY
done
ADER_B200_TRACE=gpurun_out/trace_team.json timeout 300 python bench.py --no-period --steps 50 --warmup 5 > /dev/null 2>&1
python scripts/trace_summary.py gpurun_out/trace_team.json > gpurun_out/timeline_team.txt 2>&1; rm -f gpurun_out/trace_team.json
grep "span\|attn\|qkv\|ffn\|ln_last\|lnf" gpurun_out/timeline_team.txt
