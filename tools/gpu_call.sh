mkdir -p gpurun_out
export MMU_TIMING_ONLY=1
./build/gemm_harness 8 > gpurun_out/r2c_plain8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -o gpurun_out/r2c_cfc_train_f32x2 ./build/gemm_harness 8 > gpurun_out/r2c_ncu8.log 2>&1
./build/gemm_harness 34 > gpurun_out/r2c_plain34.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -o gpurun_out/r2c_cfc_eval_plain_f32x2 ./build/gemm_harness 34 > gpurun_out/r2c_ncu34.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
