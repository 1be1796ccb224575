set -x
GLA_DBG=1 timeout 300 python tools/stress_qr.py d 8192 24
GLA_DBG=2 timeout 300 python tools/stress_qr.py d 8192 24
GLA_DBG=4 timeout 300 python tools/stress_qr.py d 8192 24
timeout 300 python tools/stress_qr.py d 8192 24
