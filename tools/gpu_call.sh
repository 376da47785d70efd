python -m pytest tests/test_gpu_headline.py -m gpu -x -q -k "reproducible" 2>&1 | tail -2
