set -x
timeout 300 python -m pytest tests/test_qr_blocked_gpu.py tests/test_cholesky_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/time_qr.py 8192 16384
timeout 300 python tools/stress_chol_concurrent.py 8192 16 high
GLA_GEMM_DBG=1 timeout 300 python tools/stress_chol_concurrent.py 8192 16 high
GLA_GEMM_DBG=2 timeout 300 python tools/stress_chol_concurrent.py 8192 16 high
GLA_GEMM_DBG=4 timeout 300 python tools/stress_chol_concurrent.py 8192 16 high
GLA_GEMM_DBG=7 timeout 300 python tools/stress_chol_concurrent.py 8192 16 high
