#!/bin/bash
mkdir -p gpurun_out
# launch list of 3 steps (times + dram bytes)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_step.csv python scratch/one_step.py 3 > gpurun_out/r2h_ncu1.log 2>&1; echo "ncu1 rc=$?"
# full captures of the small kernels of one step
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'direct_conv|conv1_tc|attn_fused|conv_rows|roi_blend|igemm_f32|quantize' -o gpurun_out/r02_small_kernels python scratch/one_step.py 1 > gpurun_out/r2h_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/*.ncu-rep
