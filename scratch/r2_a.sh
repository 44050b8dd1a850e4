#!/bin/bash
# round 2, call A: GPU tests + the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/r2a_layers.csv > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
head -c 3000 gpurun_out/r2a_bench.json
