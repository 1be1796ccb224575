set -x
timeout 100 python tools/stress_qr.py z 2048 40 high
timeout 100 python tools/stress_qr.py z 4096 20 high
timeout 100 python tools/stress_qr.py d 2048 40 high
timeout 100 python tools/stress_qr.py d 4096 30 high
timeout 100 python tools/stress_chol_concurrent.py 2048 40 high
timeout 100 python tools/stress_chol_concurrent.py 4096 30 high
timeout 200 python tools/stress_qr.py d 8192 40
timeout 200 python tools/stress_qr.py z 8192 12
timeout 200 python tools/stress_qr.py d 16384 8
timeout 200 python tools/time_qr.py 8192 16384
timeout 200 python tools/time_zqr.py 16384
