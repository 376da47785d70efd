mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_mmbt.py -x -q > gpurun_out/r2_dropout_tests.log 2>&1; echo "exit $?"; tail -30 gpurun_out/r2_dropout_tests.log
