timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_noise.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python scripts/config_latency.py 2>&1 | tail -8
