set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --profile-csv gpurun_out/r01_layers_final2.csv > gpurun_out/r01_bench_final2.log 2> gpurun_out/r01_bench_final2.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r01_bench_reference.log 2>&1; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 110 --csv --log-file gpurun_out/r01_launches_final2.csv python bench.py --steps 3 --warmup 3 --cpu-tiles 0 > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 45 -c 15 -o gpurun_out/prof_r01_gemm2 python bench.py --steps 3 --warmup 3 --cpu-tiles 0 > gpurun_out/ncu_g2.log 2>&1; echo "ncu gemm2 rc $?"
ncu --set full --clock-control none --import-source on -k regex:"attn_fused|tc_conv|conv1_tc|conv_rows" -s 21 -c 7 -o gpurun_out/prof_r01_misc python bench.py --steps 3 --warmup 3 --cpu-tiles 0 > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc rc $?"
tail -c 1500 gpurun_out/r01_bench_final2.log
