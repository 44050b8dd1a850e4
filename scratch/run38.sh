timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; timeout 300 python bench.py --steps 8 --warmup 3 --cpu-tiles 0 "$@" > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],1), d['quality'])
"; tail -3 gpurun_out/err_$name.log; }
run ms1
run ms2
