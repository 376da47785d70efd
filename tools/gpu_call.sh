mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_mmbt.py -m gpu -x -q ) > gpurun_out/r2k_pytest_mmbt.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2k_pytest_mmbt.log
python tools/bench_mmbt.py --bert-dropout 0.1 > gpurun_out/r2k_mmbt_drop01.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r2k_mmbt_drop01.log
MMU_ATTN_UNFUSED=1 python tools/bench_mmbt.py --bert-dropout 0.1 > gpurun_out/r2k_mmbt_drop01_unfused.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r2k_mmbt_drop01_unfused.log
python tools/bench_mmbt.py > gpurun_out/r2k_mmbt_drop0.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r2k_mmbt_drop0.log
