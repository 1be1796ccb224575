set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tsqr_stream -c 1 -o gpurun_out/prof_tsqr python tools/time_tsqr.py > gpurun_out/prof_tsqr.log 2>&1
ls -la gpurun_out | head
timeout 300 python tools/time_zqr.py 8192 16384
timeout 100 python tools/time_chol.py 4096 2>&1 | head -2
