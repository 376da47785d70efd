#!/bin/bash
# Runs every case of build/gemm_harness, each under its own timeout (a protocol bug must not hang the box).
n=$(./build/gemm_harness)
for i in $(seq 0 $((n-1))); do
  timeout 60 ./build/gemm_harness $i
  echo "exit=$?"
done
