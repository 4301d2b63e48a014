#!/bin/bash
# Full continual-learning runs on the shipped splits (data_cache/ = copy of the reference's data/ dir).
set -x
R=gpurun_out/results
mkdir -p $R
run() { name=$1; shift; ( time timeout 1500 python -m ader_b200.main --data_root=data_cache --results_root $R "$@" ) > $R/$name.log 2>&1; grep -E "Average|Total time|real" $R/$name.log | tail -3; }
run diginetica_ader --dataset=DIGINETICA --save_dir=ADER
run yoochoose_ader --dataset=YOOCHOOSE --save_dir=ADER --lambda_=1.0 --batch_size=512 --test_batch=64
run diginetica_finetune --dataset=DIGINETICA --save_dir=finetune --finetune=True
run diginetica_er_random --dataset=DIGINETICA --save_dir=ER-random --selection=random --disable_distillation=True
run diginetica_ewc --dataset=DIGINETICA --save_dir=EWC --ewc=True --lambda_=500 --max_periods 6
