mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck ./build/gemm_harness 22 > gpurun_out/r2i_memcheck22.log 2>&1; echo "exit $?"; head -20 gpurun_out/r2i_memcheck22.log
