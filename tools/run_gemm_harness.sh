#!/bin/bash
# Runs every harness case in its own process under a timeout; logs to gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/gemm_harness.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
N=$(./build/gemm_harness)
for i in $(seq 0 $((N-1))); do
  timeout 120 ./build/gemm_harness $i >> $LOG 2>&1
  echo "exit=$?" >> $LOG
done
grep -E "RESULT|TIMING|exit=|error|timed out" $LOG
