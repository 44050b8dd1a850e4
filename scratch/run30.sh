CIC_PIPE_TIMELINE=1 timeout 300 python bench.py --steps 4 --warmup 3 --cpu-tiles 0 --enc-chunks 8,16,16,16,8 --dec-chunks 8,16,16,16,8 > gpurun_out/bench_tl2.log 2> gpurun_out/err_tl2.log
grep "phase timeline" gpurun_out/err_tl2.log | tail -2
