set -x
timeout 300 python -m pytest tests/test_cholesky_gpu.py tests/test_determinism_gpu.py -x -q 2>&1 | tail -3
timeout 100 python tools/time_chol.py 1024 4096 8192 2>&1 | head -3
timeout 100 python tools/time_qr.py 16384 2>&1 | head -1
