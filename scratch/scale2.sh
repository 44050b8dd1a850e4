nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r01_bench_n2.log 2> gpurun_out/r01_bench_n2.err; echo "n2 rc $?"
tail -c 1200 gpurun_out/r01_bench_n2.log; tail -5 gpurun_out/r01_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 3 > gpurun_out/r01_bench_ref_n2.log 2>&1; echo "ref n2 rc $?"; tail -c 300 gpurun_out/r01_bench_ref_n2.log
