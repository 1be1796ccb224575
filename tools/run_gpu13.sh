set -x
GLA_DBG=64 timeout 400 python tools/stress_qr.py z 8192 16
GLA_DBG=64 timeout 400 python tools/stress_qr.py d 8192 60
