python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused_batch_axis or batch_axis_attention" 2>&1 | tail -2
