#!/bin/bash
# One bounded GPU-box pass: bench line, ncu launch list, ncu --set full of the step's kernels, GPU parity tests.
# Every leg has its own timeout; outputs land in gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
T0=$(date +%s)
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs.txt
timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r1d.csv python scripts/ncu_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "launches rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs.txt
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_tc_logits|k_ffn_bwd|k_attn_bwd|k_wgrad|k_qkv_bwd|k_ffn_fwd|k_attn_ln_fwd|k_qkv_fwd' \
    -f -o gpurun_out/step_top_full python scripts/ncu_step.py 1 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs.txt
timeout 150 ncu --profile-from-start off --set full --clock-control none \
    -f -o gpurun_out/step_all_full python scripts/ncu_step.py 1 > gpurun_out/ncu_all.log 2>&1
echo "all rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs.txt
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs.txt
for r in step_top_full step_all_full; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
du -sm gpurun_out/* | sort -n | tail -5
SZ=$(du -sm gpurun_out | cut -f1)
if [ "$SZ" -gt 58 ]; then rm -f gpurun_out/step_all_full.ncu-rep; fi
SZ=$(du -sm gpurun_out | cut -f1)
if [ "$SZ" -gt 58 ]; then rm -f gpurun_out/step_top_full.ncu-rep; fi
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/legs.txt
head -c 600 gpurun_out/bench_n1.json
