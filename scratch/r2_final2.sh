#!/bin/bash
# round 2 final evidence, part 2: ncu --set full of the kernels DESIGN.md quotes (kept under 64 MiB: no source import, few launches)
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'tc_deconv2_kernel|tc_conv_kernel|strip2|conv_rows' -o gpurun_out/r02_decoder_kernels python scratch/one_step.py 1 > gpurun_out/r02_ncu_full1.log 2>&1; echo "ncu full1 rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'tc_gemm2_kernel' -c 8 -o gpurun_out/r02_gemm2_kernels python scratch/one_step.py 1 > gpurun_out/r02_ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
ls -la gpurun_out; du -sh gpurun_out
