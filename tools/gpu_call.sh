mkdir -p gpurun_out
for i in 1 2; do
python bench.py --no-incumbent --no-cpu --no-other-configs > gpurun_out/r2p_bench$i.json 2> gpurun_out/r2p_bench$i.err; echo "bench exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/r2p_bench$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['config'].get('legs'))"
done
