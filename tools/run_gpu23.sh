set -x
timeout 300 python tools/stress_qr.py d 2048 40 high
timeout 300 python tools/stress_qr.py d 4096 40 high
timeout 300 python tools/stress_qr.py z 2048 30 high
