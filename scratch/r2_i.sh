#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/bench_metrics.py > gpurun_out/r2i_metrics.log 2>&1; cat gpurun_out/r2i_metrics.log
timeout 300 python scratch/bench_rd.py > gpurun_out/r2i_rd.log 2>&1; cat gpurun_out/r2i_rd.log
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "metrics" 2>&1 | tail -3
