#!/bin/bash
# ncu evidence for round 1: launch list of one bench step + full captures of the top kernels.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --profile"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 30 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 3 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ce_uncertainty -s 2 -c 2 -o gpurun_out/prof_epi $CMD > gpurun_out/ncu_epi.log 2>&1
ls -la gpurun_out/
