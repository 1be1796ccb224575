set -x
GLA_DBG=128 timeout 400 python tools/stress_qr.py z 8192 12
GLA_DBG=256 timeout 400 python tools/stress_qr.py z 8192 12
