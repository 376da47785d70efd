mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "headline8 exit $?"
$TR --master-port 29522 bench.py --gpus 8 --workload sweep10 --steps 977 --warmup 3 > gpurun_out/r2_sweep10_8gpu.json 2> gpurun_out/r2_sweep10_8gpu.err; echo "sweep10 exit $?"
$TR --master-port 29523 bench.py --gpus 8 --workload sweep43 --steps 40 --warmup 3 > gpurun_out/r2_sweep43_8gpu.json 2> gpurun_out/r2_sweep43_8gpu.err; echo "sweep43 exit $?"
$TR --master-port 29524 bench.py --gpus 8 --live-tokens --steps 20 --warmup 3 > gpurun_out/r2_live_8gpu.json 2> gpurun_out/r2_live_8gpu.err; echo "live8 exit $?"
