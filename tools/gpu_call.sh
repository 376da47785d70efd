mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2f_pytest_gpu.log
python tools/step_timeline.py 5 > gpurun_out/r2f_timeline.json 2> gpurun_out/r2f_timeline.err; echo "exit $?"; tail -1 gpurun_out/r2f_timeline.err
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"; head -c 300 gpurun_out/r2f_bench.json; echo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2f_smoke.log
