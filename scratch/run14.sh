timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() { name=$1; chunks=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --cpu-tiles 0 --e2e-chunks $chunks --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo "$name rc $?"; grep -E "deconv4|rd/conv2" gpurun_out/layers_$name.csv | cut -d, -f1,2 | tr '\n' ' '; echo; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('   value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'layers', round(d['layers_ms_per_step'],3), d['clocks'])
"; tail -2 gpurun_out/err_$name.log; }
run mg1 auto A=1
run mg0 auto CIC_TC_MERGE=0
run mg1nw2 auto CIC_TC_NW=2
