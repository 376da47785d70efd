mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --profile"
$CMD > gpurun_out/r2q_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:battn_fwd -s 3 -c 2 -o gpurun_out/r2q_battn $CMD > gpurun_out/r2q_ncu_battn.log 2>&1
echo "exit $?"; tail -3 gpurun_out/r2q_ncu_battn.log
