mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2b_pytest_gpu.log
python tools/e2e_probe.py > gpurun_out/r2_e2e_probe_a.json 2> gpurun_out/r2_e2e_probe_a.err; echo "probe exit $?"; cat gpurun_out/r2_e2e_probe_a.json | head -c 900; grep -m3 "Error\|error" gpurun_out/r2_e2e_probe_a.err
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"; head -c 1500 gpurun_out/r2b_bench.json
