timeout 120 python -m pytest tests/test_gpu_kernels.py tests/test_noise.py -m gpu -q -x 2>&1 | tail -3
timeout 120 python scripts/phase_prof.py 8 2>&1 | tail -11
timeout 200 python scripts/quick_bench.py 2>&1 | tail -3
FHE_PBS_DBG=32 timeout 200 python scripts/quick_bench.py p8 2>&1 | tail -1
