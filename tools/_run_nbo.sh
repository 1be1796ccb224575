set -x
GLA_CUDA_LIB=$PWD/tools/libgla_nbo384.so timeout 100 python tools/time_qr.py 4096 8192 16384 2>&1 | grep "n=\|gram"
GLA_CUDA_LIB=$PWD/tools/libgla_nbo512.so timeout 100 python tools/time_qr.py 4096 8192 16384 2>&1 | grep "n=\|gram"
