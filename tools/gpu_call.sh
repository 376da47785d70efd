mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dropout.py -m gpu -x -q ) > gpurun_out/r2m_pytest_mmbt.log 2>&1; echo "pytest exit $?"; tail -1 gpurun_out/r2m_pytest_mmbt.log
python tools/bench_mmbt.py --bert-dropout 0.1 --no-cpu > gpurun_out/r2m_mmbt_drop01.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2m_mmbt_drop01.log | head -c 300; echo
python tools/bench_mmbt.py --no-cpu > gpurun_out/r2m_mmbt_drop0.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2m_mmbt_drop0.log | head -c 300; echo
