#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py tests/test_gpu_full_size.py -m gpu -x -q -k "metrics or rans or saliency_mask or ms_ssim or ragged or u8 or stream" > gpurun_out/r2k_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -15 gpurun_out/r2k_memcheck.log
