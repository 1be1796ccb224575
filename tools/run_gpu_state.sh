# One gpurun call: smoke, GPU parity tests, bench line (+ reference arm), sanitizer on the batched kernel, ncu launch
# lists and full captures of the top kernels.  FULL=1 adds the n=16384 QR launch list and the DMMA GEMM capture
# (several GPU-minutes).  The torch-free harness tools/sweep_batched must have been built (see its header).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1200 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
head -c 400 gpurun_out/bench_reference.json
for tool in memcheck racecheck; do
  for b in 1 1001 4098; do
    timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 tools/sweep_batched $b 1 2>&1 | tail -2
  done
done > gpurun_out/sanitizer_batched.txt 2>&1
cat gpurun_out/sanitizer_batched.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-other --skip-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:batched_qr32 -s 2 -c 1 -f -o gpurun_out/prof_batched tools/sweep_batched 1048576 1 > gpurun_out/prof_batched.log 2>&1
if [ -n "$FULL" ]; then
  timeout 300 python tools/time_qr.py 1024 4096 8192 16384 > gpurun_out/time_qr.txt 2>&1
  cat gpurun_out/time_qr.txt
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_qr16384.csv python tools/prof_qr.py 16384 > gpurun_out/prof_qr.log 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:gemm_tn_dmma_kernel -s 400 -c 2 -f -o gpurun_out/prof_gemm python tools/prof_qr.py 16384 > gpurun_out/prof_gemm.log 2>&1
fi
ls -la gpurun_out
