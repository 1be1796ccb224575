set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 200 python tools/time_tsqr.py
timeout 300 python tools/time_zqr.py 2048 4096 8192
GLA_ZGEMM_FMA=1 timeout 300 python tools/time_zqr.py 4096
