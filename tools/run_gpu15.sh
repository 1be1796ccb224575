set -x
GLA_DBG=512 timeout 200 python tools/stress_qr.py z 8192 12
GLA_DBG=513 timeout 200 python tools/stress_qr.py z 8192 12
