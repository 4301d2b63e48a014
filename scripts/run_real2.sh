#!/bin/bash
R=gpurun_out/results
mkdir -p $R
run() { name=$1; shift; ( time timeout 1500 python -m ader_b200.main --data_root=data_cache --results_root $R "$@" ) > $R/$name.log 2>&1; echo $name; grep -E "Average|Total time|real" $R/$name.log | tail -3; }
run er_herding --dataset=DIGINETICA --disable_distillation=True --save_dir=ER-herding
run er_loss --dataset=DIGINETICA --disable_distillation=True --selection=loss --save_dir=ER-loss
run equal --dataset=DIGINETICA --equal_exemplar=True --save_dir=equal_exemplar
run fix --dataset=DIGINETICA --fix_lambda=True --save_dir=fix_lambda
run dropout --dataset=DIGINETICA --dropout=True --save_dir=dropout
run exemplar20k --dataset=DIGINETICA --exemplar_size=20000 --save_dir=exemplar20k
run joint --dataset=DIGINETICA --joint=True --save_dir=joint
