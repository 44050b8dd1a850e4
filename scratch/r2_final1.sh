#!/bin/bash
# round 2 final evidence, part 1: tests, smoke, bench (N=1), reference arm, ncu launch lists (small outputs only)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_gpu_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/r02_layers.csv > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1.err
SECONDS=0; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$? in ${SECONDS}s"; head -c 900 gpurun_out/r02_bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra-configs --cpu-tiles 0 > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_step.csv python scratch/one_step.py 3 > gpurun_out/r02_ncu_step.log 2>&1; echo "ncu step rc=$?"
du -sh gpurun_out
