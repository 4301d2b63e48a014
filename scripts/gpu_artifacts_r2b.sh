#!/bin/bash
# Final state of round 2 (second session): full GPU test suite, the bench line, ncu launch list + --set full pages of the
# kernels this session changed, compute-sanitizer memcheck over the fused-step tests.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2b_pytest_final.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/r2b_pytest_final.log
timeout 240 python bench.py > gpurun_out/r2b_bench_final.json 2> gpurun_out/r2b_bench_final.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))"
ADER_B200_TRACE=gpurun_out/r2b_trace.json timeout 100 python bench.py --no-period > /dev/null 2>&1
python scripts/trace_summary.py gpurun_out/r2b_trace.json > gpurun_out/r2b_timeline_final.txt 2>&1; rm -f gpurun_out/r2b_trace.json
timeout 100 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2b_launches_step.csv python scripts/ncu_step.py 2 > gpurun_out/r2b_ncu_launches.log 2>&1
echo "launches rc=$? t=$(( $(date +%s) - T0 ))"
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_tc2|k_adam|k_scatter_apply|k_wgrad2|k_lnf_bwd_drep|k_attn_bwd_s1|k_qkv_bwd' \
    -f -o gpurun_out/r2b_step_full python scripts/ncu_step.py 1 > gpurun_out/r2b_ncu_step_full.log 2>&1
echo "step full rc=$? t=$(( $(date +%s) - T0 ))"
[ -f gpurun_out/r2b_step_full.ncu-rep ] && ncu -i gpurun_out/r2b_step_full.ncu-rep --page raw --csv > gpurun_out/r2b_step_full.raw.csv 2>/dev/null
rm -f gpurun_out/r2b_step_full.ncu-rep
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_encoder_fused.py -m gpu -x -q \
    -k "train_dag_entry or fused_train_step or fused_gradients" > gpurun_out/r2b_sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$? t=$(( $(date +%s) - T0 ))"; tail -4 gpurun_out/r2b_sanitizer_memcheck.log
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench_final.json"))
print(round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["roofline"]["frac"], d["gpu_launches_per_step"], d["clocks"])
print([ (p["period"], round(p["sessions_per_s"]), round(p["steady_sessions_per_s"]), round(p["eval_rows_per_s"])) for p in d["period_metric"]["periods"]])
PY
