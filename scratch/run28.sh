run() { name=$1; shift; timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 "$@" > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],1), d['quality']['psnr_db'])
"; tail -3 gpurun_out/err_$name.log; }
run pipe --e2e-mode pipelined
run ph_default --e2e-mode phased
run ph_a --e2e-mode phased --enc-chunks 8,16,40 --dec-chunks 40,16,8
run ph_b --e2e-mode phased --enc-chunks 8,8,16,32 --dec-chunks 24,24,8,8
run ph_c --e2e-mode phased --enc-chunks 4,4,8,16,32 --dec-chunks 32,16,8,4,4
run ph_d --e2e-mode phased --enc-chunks 8,24,32 --dec-chunks 32,24,8
run ph_default2 --e2e-mode phased
