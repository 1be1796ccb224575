set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_qr_blocked_gpu.py tests/test_determinism_gpu.py -x -q 2>&1 | tail -3
timeout 100 python tools/time_qr.py 1024 4096 16384
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_chol4096.csv python tools/time_chol.py 4096 > gpurun_out/chol.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_qr1024.csv python tools/prof_qr.py 1024 1 > gpurun_out/qr1024.log 2>&1
