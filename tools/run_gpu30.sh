set -x
GLA_ZGEMM_FMA=1 timeout 200 python tools/stress_qr.py z 2048 30 high
GLA_QR_NO_OVERLAP=1 timeout 100 python tools/stress_qr.py z 2048 30 high
