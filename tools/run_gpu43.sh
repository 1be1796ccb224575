set -x
GLA_BATCHED_VARIANT=4 timeout 200 python -m pytest tests/test_batched_qr_gpu.py -x -q 2>&1 | tail -2
python tools/time_batched.py 2>&1 | head -1
GLA_BATCHED_VARIANT=4 python tools/time_batched.py 2>&1 | head -1
GLA_BATCHED_VARIANT=6 python tools/time_batched.py 2>&1 | head -1
GLA_BATCHED_VARIANT=5 python tools/time_batched.py 2>&1 | head -1
