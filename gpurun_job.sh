timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generations" 2>&1 | grep -E "Error|error|assert|passed|failed" | head -20
