#!/bin/bash
# ncu evidence for the MMBT path: launch lists of one train step + eval forwards (pooled tokens and
# raw images) and full captures of the fused attention kernel and the top image-encoder kernels.
set -x
mkdir -p gpurun_out
CMD="python tools/bench_mmbt.py --steps 1 --no-cpu"
$CMD > gpurun_out/mmbt_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/mmbt_launches_tokens.csv $CMD > gpurun_out/ncu_mmbt_list.log 2>&1
$CMD > gpurun_out/mmbt_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fattn_fwd -s 13 -c 2 -o gpurun_out/prof_fattn $CMD > gpurun_out/ncu_fattn.log 2>&1
$CMD --images > gpurun_out/mmbt_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/mmbt_launches_images.csv $CMD --images > gpurun_out/ncu_mmbt_list_img.log 2>&1
ls -la gpurun_out/ | tail -8
