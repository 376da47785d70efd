mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-other-configs --host-dtype bf16 > gpurun_out/r2_bench_8gpu_bf16host.json 2> gpurun_out/r2_bench_8gpu_bf16host.err; echo "headline8 bf16 host exit $?"
$TR --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --no-other-configs --live-tokens --host-dtype bf16 > gpurun_out/r2_live_8gpu_bf16host.json 2> gpurun_out/r2_live_8gpu_bf16host.err; echo "live8 bf16 host exit $?"
