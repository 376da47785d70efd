#!/bin/bash
# ncu evidence for the MMBT path: launch lists of one train step + eval forwards (pooled tokens, raw
# images, image encoder alone) and full captures of the fused attention kernels.
set -x
mkdir -p gpurun_out
CMD="python tools/bench_mmbt.py --steps 1 --no-cpu"
$CMD > gpurun_out/mmbt_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/mmbt_launches_tokens.csv $CMD > gpurun_out/ncu_mmbt_list.log 2>&1
$CMD > gpurun_out/mmbt_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fattn_fwd -s 13 -c 2 -o gpurun_out/prof_fattn $CMD > gpurun_out/ncu_fattn.log 2>&1
$CMD > gpurun_out/mmbt_plain2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fattn_bwd -s 13 -c 2 -o gpurun_out/prof_fattn_bwd $CMD > gpurun_out/ncu_fattn_bwd.log 2>&1
python tools/bench_imgenc.py --steps 1 > gpurun_out/imgenc_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/imgenc_launches_final.csv python tools/bench_imgenc.py --steps 1 > gpurun_out/ncu_imgenc_final.log 2>&1
ls -la gpurun_out/ | tail -6
