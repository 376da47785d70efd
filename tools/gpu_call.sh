mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke exit $?"; grep -c "smoke\[" gpurun_out/r2z_smoke.log
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['eval_gemms']['frac'], d['whole_step']['frac'], d['gpu_launches'])"
