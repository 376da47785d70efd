# The command this repo hands to gpurun for a full verification of the tree (tests, smoke, bench):
#   gpurun --timeout 400 -- 'bash tools/gpu_call.sh'
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep -c "smoke\[" gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench.json
