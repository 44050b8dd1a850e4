set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
B="python bench.py --steps 3 --warmup 3 --cpu-tiles 0"
python bench.py --profile-csv gpurun_out/r01_layers_final6.csv > gpurun_out/r01_bench_final6.log 2> gpurun_out/r01_bench_final6.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference6.log 2>&1; echo "ref rc $?"
python bench.py --ssim exact --cpu-tiles 0 > gpurun_out/r01_bench_final6_exact_ssim.log 2>&1; echo "exact rc $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 247 -c 70 --csv --log-file gpurun_out/r01_launches_final6.csv $B > gpurun_out/ncu_l6.log 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:metrics_f32_fast -s 3 -c 1 -o gpurun_out/prof_r01h_metrics $B > gpurun_out/ncu_m6.log 2>&1; echo "ncu metrics rc $?"
ncu -i gpurun_out/prof_r01h_metrics.ncu-rep --page raw --csv > gpurun_out/rawh_metrics.csv 2>/dev/null
ncu -i gpurun_out/prof_r01h_metrics.ncu-rep --page source --csv --print-source sass > gpurun_out/srch_metrics.csv 2>/dev/null; rm -f gpurun_out/prof_r01h_metrics.ncu-rep
tail -c 300 gpurun_out/r01_bench_final6.log; tail -3 gpurun_out/r01_bench_final6.err
