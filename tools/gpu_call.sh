mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_suite3.log 2>&1; echo "suite exit $?"; tail -12 gpurun_out/r2_gpu_suite3.log
python tools/attn_probe.py > gpurun_out/r2_attn_probe2.log 2>&1; cat gpurun_out/r2_attn_probe2.log
python bench.py --no-cpu --no-incumbent --no-other-configs > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench exit $?"
