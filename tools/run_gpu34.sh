set -x
python tools/time_e2e.py
GLA_BATCH_CHUNK_MB=16 GLA_BATCH_STREAMS=4 python tools/time_e2e.py 2>&1 | head -1
GLA_BATCH_CHUNK_MB=256 GLA_BATCH_STREAMS=3 python tools/time_e2e.py 2>&1 | head -1
GLA_BATCH_CHUNK_MB=128 GLA_BATCH_STREAMS=6 python tools/time_e2e.py 2>&1 | head -1
timeout 100 python tools/time_chol.py 4096 2>&1 | head -1
timeout 100 python tools/time_qr.py 1024 2>&1 | head -1
