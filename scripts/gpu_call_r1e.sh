#!/bin/bash
# Second bounded GPU pass: parity tests first (DAG entry), bench with the DAG and with the three-group sequence,
# launch list, ncu --set full of the kernels the first pass did not capture.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_e.txt
timeout 150 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_dag.json 2> gpurun_out/bench_dag.err
echo "bench dag rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_e.txt
ADER_B200_STEP_IMPL=groups timeout 150 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_groups.json 2> gpurun_out/bench_groups.err
echo "bench groups rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_e.txt
timeout 100 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r1e.csv python scripts/ncu_step.py 2 > gpurun_out/ncu_launches_e.log 2>&1
echo "launches rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_e.txt
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_scatter_small|k_pos_grad|k_adam|k_reduce|k_teacher|k_merge|k_lnf|k_ln_param|k_ln_last|k_pack_tiles|k_gather' \
    -f -o gpurun_out/step_rest_full python scripts/ncu_step.py 1 > gpurun_out/ncu_rest.log 2>&1
echo "rest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_e.txt
[ -f gpurun_out/step_rest_full.ncu-rep ] && ncu -i gpurun_out/step_rest_full.ncu-rep --page raw --csv > gpurun_out/step_rest_full.raw.csv 2>/dev/null
rm -f gpurun_out/step_top_full.ncu-rep
tail -5 gpurun_out/pytest_gpu_e.log
cat gpurun_out/legs_e.txt
head -c 300 gpurun_out/bench_dag.json; echo
head -c 300 gpurun_out/bench_groups.json; echo
tail -3 gpurun_out/bench_dag.err
