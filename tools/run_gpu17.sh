set -x
GLA_QR_NO_OVERLAP=1 timeout 400 python tools/stress_qr.py d 16384 10
GLA_DBG=64 timeout 400 python tools/stress_qr.py d 16384 10
timeout 400 python tools/stress_qr.py d 16384 10
