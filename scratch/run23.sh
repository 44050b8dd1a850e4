CIC_PIPE_TIMELINE=1 timeout 300 python bench.py --steps 4 --warmup 3 --cpu-tiles 0 > gpurun_out/bench_tl.log 2> gpurun_out/err_tl.log
grep "pipe timeline" gpurun_out/err_tl.log | tail -3
for ch in "8,16,32,8" "4,12,16,24,8" "8,24,24,8" "6,10,16,24,8" "8,16,36,4"; do
echo -n "$ch: "; timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 --e2e-chunks $ch 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
