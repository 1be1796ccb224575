set -x
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_tsqr_gpu.py -x -q -k "777 or 10-4 or 257 or 65-64" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_cholesky_gpu.py -x -q -k "50" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_batched_qr_gpu.py -x -q 2>&1 | tail -6
