nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r01_bench_n8.log 2> gpurun_out/r01_bench_n8.err; echo "n4 rc $?"
tail -c 300 gpurun_out/r01_bench_n8.log; tail -3 gpurun_out/r01_bench_n8.err
