set -x
timeout 100 python tools/time_chol.py 4096 2>&1 | head -8
