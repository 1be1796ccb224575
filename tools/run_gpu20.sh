set -x
GLA_QR_OVERLAP=1 timeout 300 python tools/stress_qr.py z 8192 16
GLA_QR_OVERLAP=1 timeout 300 python tools/stress_qr.py d 16384 12
GLA_QR_OVERLAP=1 timeout 300 python tools/stress_qr.py d 8192 80
GLA_QR_OVERLAP=1 timeout 300 python tools/time_qr.py 8192 16384
