set -x
timeout 400 python tools/stress_qr.py z 8192 16
timeout 400 python tools/stress_qr.py d 8192 100
timeout 300 python tools/time_qr.py 16384
