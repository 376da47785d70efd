mkdir -p gpurun_out
for v in base same zdir; do
  for c in 7 8 9 10 11 19; do
    MMU_FORCE_TIMING=1 timeout 120 ./build/gemm_harness_$v $c 2>&1 | grep -E "RESULT|TIMING|error|Error" | sed "s/^/[$v] /"
  done
done > gpurun_out/r2_harness_variants.log 2>&1
cat gpurun_out/r2_harness_variants.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_suite.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/r2_gpu_suite.log
