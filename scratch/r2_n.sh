#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'conv_rows' -o gpurun_out/r02_gen_tail python scratch/one_step.py 1 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
