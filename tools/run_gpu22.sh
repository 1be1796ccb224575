set -x
timeout 100 python tools/time_chol.py 4096 2>&1 | head -2
timeout 300 python tools/stress_chol_concurrent.py 4096 40 high
timeout 300 python tools/stress_chol_concurrent.py 4096 30 none
timeout 300 python tools/stress_chol_concurrent.py 2048 40 high
