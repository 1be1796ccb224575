set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tsqr_gpu.py tests/test_qr_blocked_gpu.py -x -q 2>&1 | tail -8
timeout 200 python tools/time_tsqr.py
timeout 200 python tools/time_qr.py 8192 16384
