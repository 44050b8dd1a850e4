set -x
B="python bench.py --steps 3 --warmup 3 --cpu-tiles 0"
python bench.py --steps 10 --warmup 3 --profile-csv gpurun_out/r01_layers_final3.csv > gpurun_out/r01_bench_final3.log 2> gpurun_out/r01_bench_final3.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference3.log 2>&1; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 247 -c 70 --csv --log-file gpurun_out/r01_launches_final3.csv $B > gpurun_out/ncu_l3.log 2>&1; echo "ncu list rc $?"
cap() { name=$1; shift; ncu --set full --clock-control none --import-source on "$@" -o gpurun_out/prof_r01f_$name $B > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc $?";
  ncu -i gpurun_out/prof_r01f_$name.ncu-rep --page raw --csv > gpurun_out/rawf_$name.csv 2>/dev/null;
  ncu -i gpurun_out/prof_r01f_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/srcf_$name.csv 2>/dev/null; rm -f gpurun_out/prof_r01f_$name.ncu-rep; }
cap deconv4 -k regex:tc_conv_kernel -s 10 -c 1
cap conv3 -k regex:tc_gemm2 -s 46 -c 1
cap deconv2 -k regex:tc_gemm2 -s 56 -c 1
cap metrics -k regex:metrics_f32_packed -s 3 -c 1
du -sm gpurun_out
tail -c 600 gpurun_out/r01_bench_final3.log
