set -x
GLA_BATCHED_VARIANT=4 timeout 200 python -m pytest tests/test_batched_qr_gpu.py -x -q 2>&1 | tail -2
python tools/time_batched.py 2>&1 | head -1
GLA_BATCHED_VARIANT=4 python tools/time_batched.py 2>&1 | head -1
GLA_BATCHED_VARIANT=6 python tools/time_batched.py 2>&1 | head -1
GLA_BATCHED_VARIANT=5 python tools/time_batched.py 2>&1 | head -1
timeout 300 python -m pytest tests/test_cholesky_gpu.py tests/test_qr_blocked_gpu.py tests/test_determinism_gpu.py -x -q 2>&1 | tail -2
timeout 100 python tools/time_chol.py 1024 4096 2>&1 | grep "chol n"
timeout 100 python tools/time_qr.py 1024 4096 16384 2>&1 | grep "n="
