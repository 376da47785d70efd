mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench exit $?"
python bench.py --live-tokens --no-cpu --no-incumbent --no-other-configs > gpurun_out/r2_bench3_live.json 2> gpurun_out/r2_bench3_live.err; echo "live exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench3_ref.json 2> gpurun_out/r2_bench3_ref.err; echo "ref exit $?"
python tools/bench_mmbt.py --no-cpu > gpurun_out/r2_mmbt_drop0.log 2>&1; tail -3 gpurun_out/r2_mmbt_drop0.log
python tools/bench_mmbt.py --no-cpu --bert-dropout 0.1 > gpurun_out/r2_mmbt_drop01.log 2>&1; tail -3 gpurun_out/r2_mmbt_drop01.log
python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r2_prof_plain2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches2.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r2_prof_ncu2.log 2>&1; echo "launch list exit $?"
ATTN_ONCE=1 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 2 -o gpurun_out/r2_attn_fwd_pm python tools/attn_probe.py > gpurun_out/r2_attn_ncu2.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r2_attn_fwd_pm.ncu-rep --page raw --csv > gpurun_out/r2_attn_fwd_pm_raw.csv 2>/dev/null
for c in 7 8 10; do ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 1 -o gpurun_out/r2_gemm_case$c ./build/gemm_harness $c > /dev/null 2>&1; ncu -i gpurun_out/r2_gemm_case$c.ncu-rep --page raw --csv > gpurun_out/r2_gemm_case${c}_raw.csv 2>/dev/null; done; echo captures done
