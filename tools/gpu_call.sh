mkdir -p gpurun_out
T0=$(date +%s)
{
for i in 36 37 38 39; do timeout 60 ./build/gemm_harness $i | grep -E "RESULT|mismatch"; done
for i in 40 41; do MMU_TIMING_ONLY=1 timeout 60 ./build/gemm_harness $i | grep TIMING; done
} > gpurun_out/r2v_harness_erf.log 2>&1; cat gpurun_out/r2v_harness_erf.log
echo "t=$(( $(date +%s) - T0 ))"
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2v_pytest_gpu.log
cp gpurun_out/measured_errors.json gpurun_out/r2v_measured_errors.json 2>/dev/null
echo "t=$(( $(date +%s) - T0 ))"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke exit $?"; grep -c "smoke\[" gpurun_out/r2v_smoke.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['eval_gemms']['frac'], d['whole_step']['frac'], d['gpu_launches'])"
echo "t=$(( $(date +%s) - T0 ))"
timeout 100 python tools/bench_mmbt.py --no-cpu > gpurun_out/r2v_mmbt.json 2>&1; tail -1 gpurun_out/r2v_mmbt.json | cut -c1-400
echo "t=$(( $(date +%s) - T0 ))"
