set -x
GLA_QR_NO_OVERLAP=1 timeout 300 python tools/stress_qr.py z 8192 8
GLA_QR_NO_OVERLAP=1 timeout 300 python tools/stress_qr.py d 8192 16
timeout 300 python tools/stress_qr.py d 8192 16
timeout 200 python tools/time_chol.py 4096 2>&1 | head -3
timeout 300 python -m pytest tests/test_cholesky_gpu.py -x -q 2>&1 | tail -3
