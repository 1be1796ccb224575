set -x
timeout 100 python tools/stress_qr.py z 4096 20 high
timeout 100 python tools/stress_qr.py z 2048 24 high
