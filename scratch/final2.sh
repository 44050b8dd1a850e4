set -x
B="python bench.py --steps 3 --warmup 3 --cpu-tiles 0"
python bench.py --steps 10 --warmup 3 --profile-csv gpurun_out/r01_layers_final2.csv > gpurun_out/r01_bench_final2.log 2> gpurun_out/r01_bench_final2.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference.log 2>&1; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 110 --csv --log-file gpurun_out/r01_launches_final2.csv $B > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc $?"
cap() { name=$1; shift; ncu --set full --clock-control none --import-source on "$@" -o gpurun_out/prof_r01_$name $B > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc $?";
  ncu -i gpurun_out/prof_r01_$name.ncu-rep --page raw --csv > gpurun_out/raw_r01_$name.csv 2>/dev/null; }
cap deconv4 -k regex:tc_conv_kernel -s 10 -c 1
cap conv3 -k regex:tc_gemm2 -s 46 -c 1
cap deconv3 -k regex:tc_gemm_kernel -s 10 -c 1
cap deconv2 -k regex:tc_gemm2 -s 56 -c 1
cap attn -k regex:attn_fused -s 3 -c 1
cap conv1 -k regex:conv1_tc -s 3 -c 1
cap convout -k regex:conv_rows -s 6 -c 1
cap bw -k regex:"roi_blend_c3|metrics_f32_packed|quantize_kernel" -s 9 -c 3
ls -l gpurun_out/*.ncu-rep
# keep the pull under the 64 MiB limit: drop the least important reports first
for f in bw convout conv1 attn deconv2; do
  sz=$(du -sm gpurun_out | cut -f1); if [ "$sz" -gt 55 ]; then rm -f gpurun_out/prof_r01_$f.ncu-rep; fi
done
du -sm gpurun_out
tail -c 600 gpurun_out/r01_bench_final2.log
