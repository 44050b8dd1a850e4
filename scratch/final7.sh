set -x
python bench.py --profile-csv gpurun_out/r01_layers_final7.csv > gpurun_out/r01_bench_final7.log 2> gpurun_out/r01_bench_final7.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference7.log 2>&1; echo "ref rc $?"
tail -c 200 gpurun_out/r01_bench_final7.log; tail -3 gpurun_out/r01_bench_final7.err
