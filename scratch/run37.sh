timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; timeout 300 python bench.py --steps 8 --warmup 3 --cpu-tiles 0 --profile-csv gpurun_out/layers_$name.csv "$@" > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; grep -E "rd/conv1" gpurun_out/layers_$name.csv | cut -d, -f2 | tr '\n' ' '; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],1))
"; tail -3 gpurun_out/err_$name.log; }
run dcb1
run dcb2
