mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2d_pytest_gpu.log
python tools/step_timeline.py 5 > gpurun_out/r2d_timeline.json 2> gpurun_out/r2d_timeline.err; echo "exit $?"; tail -1 gpurun_out/r2d_timeline.err
python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"; head -c 400 gpurun_out/r2d_bench.json; echo
