#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -u scratch/stream_timeline.py u8 > gpurun_out/r2d_tl_u8.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r2d_tl_u8.log
timeout 300 python -u scratch/stream_timeline.py f32 > gpurun_out/r2d_tl_f32.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r2d_tl_f32.log
