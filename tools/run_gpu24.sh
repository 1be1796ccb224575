set -x
GLA_DGEMM_FMA=1 timeout 300 python tools/stress_qr.py d 4096 30 high
GLA_QR_NO_OVERLAP=1 timeout 300 python tools/stress_qr.py d 4096 30 high
timeout 300 python tools/stress_qr.py d 4096 30 normal
