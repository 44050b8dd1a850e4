#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "ops tests rc=$?"; tail -15 gpurun_out/r2b_tests.log
timeout 300 python scratch/bench_metrics.py > gpurun_out/r2b_metrics.log 2>&1; cat gpurun_out/r2b_metrics.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:metrics_f32_strip -c 1 -o gpurun_out/r2b_strip python scratch/bench_metrics.py 0 > gpurun_out/r2b_ncu.log 2>&1; tail -3 gpurun_out/r2b_ncu.log
ls -la gpurun_out/*.ncu-rep
