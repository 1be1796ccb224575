set -x
GLA_DBG=16 timeout 400 python tools/stress_qr.py d 8192 20
GLA_DBG=32 timeout 300 python tools/stress_qr.py d 8192 24
timeout 300 python tools/stress_qr.py d 4096 40
