set -x
timeout 300 python -m pytest tests/test_qr_blocked_gpu.py tests/test_determinism_gpu.py -x -q 2>&1 | tail -3
timeout 100 python tools/time_qr.py 1024 4096 8192 16384
