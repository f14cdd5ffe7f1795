timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/config_latency.py 2>&1 | tail -8
