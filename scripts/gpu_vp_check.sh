#!/bin/bash
# BASELINE config 5 (1 M items, batch 4096 + 1024 KD rows): vocab-parallel logits + CE + KD fwd+bwd at N GPUs
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  scripts/bench_vocab_parallel.py > gpurun_out/vp_n$N.json 2> gpurun_out/vp_n$N.err; echo "vp N=$N rc=$?"
grep "^{" gpurun_out/vp_n$N.json | tail -1
tail -2 gpurun_out/vp_n$N.err
