set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
