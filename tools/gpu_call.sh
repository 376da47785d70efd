mkdir -p gpurun_out
export MMU_TIMING_ONLY=1
./build/gemm_harness 30 > gpurun_out/r2c_plain30.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -o gpurun_out/r2c_fold_cfc_eval_f32x2 ./build/gemm_harness 30 > gpurun_out/r2c_ncu30.log 2>&1
./build/gemm_harness 10 > gpurun_out/r2c_plain10.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -o gpurun_out/r2c_dgelu_f32x2 ./build/gemm_harness 10 > gpurun_out/r2c_ncu10.log 2>&1
tail -2 gpurun_out/r2c_ncu30.log gpurun_out/r2c_ncu10.log; ls -la gpurun_out/*.ncu-rep | tail -3
