set -x
GLA_QR_NO_OVERLAP=1 timeout 400 python tools/stress_qr.py d 8192 120
GLA_QR_NO_OVERLAP=1 timeout 400 python tools/stress_qr.py z 8192 30
GLA_ZGEMM_FMA=1 timeout 400 python tools/stress_qr.py z 8192 20
GLA_DBG=16 timeout 400 python tools/stress_qr.py d 8192 60
