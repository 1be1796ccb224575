set -x
GLA_GEMM_DBG=1 timeout 100 python tools/stress_qr.py d 4096 24 high
GLA_GEMM_DBG=2 timeout 100 python tools/stress_qr.py d 4096 24 high
GLA_GEMM_DBG=4 timeout 100 python tools/stress_qr.py d 4096 24 high
GLA_GEMM_DBG=8 timeout 100 python tools/stress_qr.py d 4096 24 high
GLA_GEMM_DBG=32 timeout 100 python tools/stress_qr.py d 4096 24 high
