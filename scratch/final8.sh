set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --profile-csv gpurun_out/r01_layers_final8.csv > gpurun_out/r01_bench_final8.log 2> gpurun_out/r01_bench_final8.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference8.log 2>&1; echo "ref rc $?"
tail -c 200 gpurun_out/r01_bench_final8.log; tail -3 gpurun_out/r01_bench_final8.err
