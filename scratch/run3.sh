python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --cpu-tiles 0 --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2>&1; echo "$name rc $?"; grep -E "hq_gen/deconv|hq_enc/conv2|rd/conv2" gpurun_out/layers_$name.csv | cut -d, -f1,2 | tr '\n' ' '; echo; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('   value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'layers', round(d['layers_ms_per_step'],3))
"; }
run def A=1
run g32k CIC_TC_BGROUP_BYTES=32768
run nw1 CIC_TC_NW=1
run nw4 CIC_TC_NW=4
