set -x
timeout 100 python tools/stress_qr.py d 4096 30 high
timeout 100 python tools/stress_qr.py d 2048 30 high
GLA_GEMM_DBG=8 timeout 100 python tools/stress_qr.py d 2048 30 high
timeout 200 python tools/stress_chol_concurrent.py 2048 40 high
