#!/bin/bash
# Measured numbers for the BASELINE.json configs bench.py does not time (bench.py = configs[1] + [2]):
# configs[0] FashionMNIST models, configs[3] MMBT (pooled tokens / raw images), the image encoder alone.
mkdir -p gpurun_out
{
  echo '{"fmnist_configs0":'; python tools/bench_fmnist.py 2>/dev/null | tail -1
  echo ',"mmbt_configs3_tokens":'; python tools/bench_mmbt.py 2>/dev/null | tail -1
  echo ',"mmbt_configs3_images":'; python tools/bench_mmbt.py --images --no-cpu 2>/dev/null | tail -1
  echo ',"image_encoder":'; python tools/bench_imgenc.py 2>/dev/null | tail -1
  echo ',"clocks":"'"$(nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader)"'"}'
} > gpurun_out/other_configs.json
python -c "import json; d=json.load(open('gpurun_out/other_configs.json')); print(json.dumps(d)[:300])"
