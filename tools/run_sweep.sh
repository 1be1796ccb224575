# One gpurun call: time batched-QR kernel variants with the torch-free harness and check each against the oracle.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v in ${VARIANTS:-0 1 2 3 4 5 6 7}; do
  GLA_BATCHED_VARIANT=$v timeout 60 tools/sweep_batched 1048576 7 gpurun_out/sweep_v$v.bin 2>&1 | tail -2
done
for v in ${ODD:-0}; do
  GLA_BATCHED_VARIANT=$v timeout 60 tools/sweep_batched 1001 1 gpurun_out/sweep_odd_v$v.bin 2>&1 | tail -1
  GLA_BATCHED_VARIANT=$v timeout 60 tools/sweep_batched 1 1 gpurun_out/sweep_one_v$v.bin 2>&1 | tail -1
  GLA_BATCHED_VARIANT=$v timeout 60 tools/sweep_batched 4098 1 gpurun_out/sweep_4098_v$v.bin 2>&1 | tail -1
done
timeout 300 python tools/sweep_check.py gpurun_out/sweep_*.bin 2>&1 | tee gpurun_out/sweep_check.txt
rm -f gpurun_out/sweep_*.bin
