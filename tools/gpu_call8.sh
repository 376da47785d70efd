mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-other-configs > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "headline8 exit $?"
