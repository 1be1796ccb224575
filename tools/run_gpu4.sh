set -x
timeout 300 python -m pytest tests/test_tsqr_gpu.py -x -q 2>&1 | tail -4
GLA_ZGEMM_FMA=1 timeout 300 python tools/time_zqr.py 6144 8192
timeout 300 python tools/time_zqr.py 6144 8192
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_chol4096.csv python tools/time_chol.py 4096 > gpurun_out/chol.log 2>&1
tail -5 gpurun_out/chol.log
