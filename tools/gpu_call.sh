mkdir -p gpurun_out
{
for c in 0 1 2 3 4 5 13 14 15 16 17 18; do timeout 120 ./build/gemm_harness_s5 $c 2>&1 | grep -E "RESULT|error|Error|timed out" | sed "s/^/[s5] /"; done
for v in s5 s4; do
  for c in 6 7 8 9 10 11 12 19 20 21; do
    timeout 120 ./build/gemm_harness_$v $c 2>&1 | grep -E "RESULT|TIMING|error|Error|timed out" | sed "s/^/[$v] /"
  done
done
} > gpurun_out/r2_harness_s5.log 2>&1
cat gpurun_out/r2_harness_s5.log
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/r2_ops_tests.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/r2_ops_tests.log
