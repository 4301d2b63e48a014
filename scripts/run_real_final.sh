#!/bin/bash
# Full continual-learning runs on the shipped splits with the final code of the round (data_cache/ = copy of the
# reference's data/ directory, git-ignored).  Prints the per-period throughput lines and the averaged metrics.
R=gpurun_out/results_final
mkdir -p $R
run() { name=$1; shift; ( time timeout 400 python -m ader_b200.main --data_root=data_cache --cache_dir=gpurun_out/pair_cache --results_root $R "$@" ) > $R/$name.log 2>&1; grep -E "train throughput|Average|Total time|real" $R/$name.log | tail -20; }
run diginetica_ader --dataset=DIGINETICA --save_dir=ADER
run yoochoose_ader --dataset=YOOCHOOSE --save_dir=ADER --lambda_=1.0 --batch_size=512 --test_batch=64
