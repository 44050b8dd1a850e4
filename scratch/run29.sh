run() { name=$1; shift; timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 "$@" > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],1))
"; tail -3 gpurun_out/err_$name.log; }
run p1 --enc-chunks 8,24,24,8 --dec-chunks 8,24,24,8
run p2 --enc-chunks 8,16,32,8 --dec-chunks 16,24,16,8
run p3 --enc-chunks 8,16,24,16 --dec-chunks 8,16,32,8
run p4 --enc-chunks 8,16,16,16,8 --dec-chunks 8,16,16,16,8
run p5 --enc-chunks 8,24,24,8 --dec-chunks 24,24,8,8
run p6 --enc-chunks 4,12,24,16,8 --dec-chunks 16,24,16,8
run pipe --e2e-mode pipelined
