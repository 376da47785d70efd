mkdir -p gpurun_out
set -x
CMD="python bench.py --steps 1 --warmup 1 --profile"
$CMD > gpurun_out/r2d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2d_launches.csv $CMD > gpurun_out/r2d_ncu_list.log 2>&1
for c in 27 28 29 30; do
./build/gemm_harness $c > gpurun_out/r2d_harness_plain_$c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/r2d_gemm_case$c ./build/gemm_harness $c > gpurun_out/r2d_ncu_case$c.log 2>&1
done
ls -la gpurun_out/ | grep r2d
