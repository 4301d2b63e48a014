#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
timeout 150 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_n2.log 2>&1
echo "pytest dist rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_n2.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$? t=$(( $(date +%s) - T0 ))" >> gpurun_out/legs_n2.txt
tail -4 gpurun_out/pytest_n2.log; cat gpurun_out/legs_n2.txt; head -c 400 gpurun_out/bench_n2.json; echo; tail -5 gpurun_out/bench_n2.err
