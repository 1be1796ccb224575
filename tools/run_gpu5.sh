set -x
timeout 200 python tools/stress_qr.py z 8192 8
GLA_ZGEMM_FMA=1 timeout 300 python tools/stress_qr.py z 8192 6
timeout 200 python tools/stress_qr.py d 8192 8
timeout 200 python tools/stress_qr.py z 3000 8
timeout 100 python tools/time_tsqr.py
GLA_TSQR_CFG=1 timeout 100 python tools/time_tsqr.py
