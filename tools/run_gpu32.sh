set -x
GLA_ZVAR=1 timeout 100 python tools/stress_qr.py z 2048 40 high
GLA_ZVAR=2 timeout 100 python tools/stress_qr.py z 2048 40 high
GLA_ZVAR=3 timeout 100 python tools/stress_qr.py z 2048 40 high
