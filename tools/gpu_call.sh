mkdir -p gpurun_out
{
for pf in 0 2 4 8 16 32; do
  for c in 9 19 20 11 21 7 8; do
    MMU_GEMM_L2PF=$pf timeout 120 ./build/gemm_harness_pf $c 2>&1 | grep -E "RESULT|TIMING|error|Error|timed out" | sed "s/^/[pf=$pf] /"
  done
done
} > gpurun_out/r2_harness_l2pf.log 2>&1
grep TIMING gpurun_out/r2_harness_l2pf.log
