timeout 120 python -m pytest tests/test_gpu_kernels.py tests/test_noise.py -m gpu -q -x 2>&1 | tail -5
timeout 120 python scripts/phase_prof.py 8 4 2>&1 | tail -22
timeout 200 python scripts/quick_bench.py 2>&1 | tail -3
