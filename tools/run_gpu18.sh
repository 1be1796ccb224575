set -x
timeout 300 python tools/stress_chol_concurrent.py 8192 12 none
timeout 300 python tools/stress_chol_concurrent.py 8192 16 normal
timeout 300 python tools/stress_chol_concurrent.py 8192 16 high
