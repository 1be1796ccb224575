set -x
GLA_DBG=32 timeout 400 python tools/stress_qr.py d 8192 60
GLA_DBG=32 timeout 400 python tools/stress_qr.py z 8192 12
GLA_DBG=32 timeout 300 python tools/time_qr.py 16384
GLA_QR_NO_OVERLAP=1 timeout 300 python tools/time_qr.py 16384
