set -x
GLA_CUDA_LIB=$PWD/tools/libgla_cuda_oldgemm.so timeout 100 python tools/stress_qr.py d 4096 24 high
timeout 100 python tools/stress_qr.py d 4096 24 high
GLA_CUDA_LIB=$PWD/tools/libgla_cuda_oldgemm.so timeout 100 python tools/stress_qr.py d 4096 24 high
timeout 100 python tools/stress_qr.py d 4096 24 high
