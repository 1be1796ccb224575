set -x
GLA_DBG=513 timeout 300 python tools/time_qr.py 4096 8192 16384
GLA_DBG=513 timeout 400 python tools/stress_qr.py d 8192 100
GLA_DBG=513 timeout 400 python tools/stress_qr.py z 8192 20
GLA_DBG=513 timeout 400 python tools/stress_qr.py d 16384 12
