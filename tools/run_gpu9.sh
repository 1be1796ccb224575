set -x
timeout 400 python tools/stress_qr.py d 8192 48
timeout 300 python tools/stress_qr.py z 8192 10
timeout 300 python tools/time_qr.py 8192 16384
