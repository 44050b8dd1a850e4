set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --profile-csv gpurun_out/r01_layers_final5.csv > gpurun_out/r01_bench_final5.log 2> gpurun_out/r01_bench_final5.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference5.log 2>&1; echo "ref rc $?"
python bench.py --e2e-mode pipelined --cpu-tiles 0 > gpurun_out/r01_bench_final5_pipelined.log 2>&1; echo "pipelined rc $?"
tail -c 300 gpurun_out/r01_bench_final5.log; tail -3 gpurun_out/r01_bench_final5.err
