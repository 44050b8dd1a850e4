set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --cpu-tiles 0"
python bench.py --steps 10 --warmup 3 --profile-csv gpurun_out/r01_layers_final4.csv > gpurun_out/r01_bench_final4.log 2> gpurun_out/r01_bench_final4.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference4.log 2>&1; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 247 -c 70 --csv --log-file gpurun_out/r01_launches_final4.csv $B > gpurun_out/ncu_l4.log 2>&1; echo "ncu list rc $?"
cap() { name=$1; shift; ncu --set full --clock-control none --import-source on "$@" -o gpurun_out/prof_r01g_$name $B > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc $?";
  ncu -i gpurun_out/prof_r01g_$name.ncu-rep --page raw --csv > gpurun_out/rawg_$name.csv 2>/dev/null; rm -f gpurun_out/prof_r01g_$name.ncu-rep; }
cap deconv4 -k regex:tc_conv_kernel -s 10 -c 1
cap conv1 -k regex:conv1_tc -s 3 -c 1
cap deconv3 -k regex:tc_gemm_kernel -s 10 -c 1
du -sm gpurun_out
tail -c 400 gpurun_out/r01_bench_final4.log
