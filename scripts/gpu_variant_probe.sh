#!/bin/bash
# the five forms of the fused step side by side (one process each, concurrently): digests must agree
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python tests/variant_probe.py > gpurun_out/vp_default.txt 2>&1 &
ADER_B200_SPLIT_ADAM=1 python tests/variant_probe.py > gpurun_out/vp_split.txt 2>&1 &
ADER_B200_FUSE_DREP=0 python tests/variant_probe.py > gpurun_out/vp_nofuse.txt 2>&1 &
ADER_B200_SCATTER=2 python tests/variant_probe.py > gpurun_out/vp_scatter2.txt 2>&1 &
ADER_B200_PACK=1 python tests/variant_probe.py > gpurun_out/vp_pack1.txt 2>&1 &
wait
tail -n 2 gpurun_out/vp_*.txt
