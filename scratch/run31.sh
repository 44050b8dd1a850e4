run() { name=$1; shift; timeout 300 python bench.py --steps 8 --warmup 3 --cpu-tiles 0 "$@" > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],1))
"; tail -3 gpurun_out/err_$name.log; }
run q1 --enc-chunks 8,16,16,16,8 --dec-chunks 8,16,16,16,8
run q2 --enc-chunks 4,12,16,16,16 --dec-chunks 16,16,16,12,4
run q3 --enc-chunks 4,12,16,32 --dec-chunks 8,24,16,12,4
run q4 --enc-chunks 4,8,20,32 --dec-chunks 12,20,16,12,4
run q5 --enc-chunks 8,16,40 --dec-chunks 8,16,16,16,8
run q6 --enc-chunks 8,24,32 --dec-chunks 8,24,24,8
run pipe --e2e-mode pipelined
