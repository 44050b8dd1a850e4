run() { name=$1; chunks=$2; shift; shift; env "$@" python bench.py --steps 5 --warmup 3 --cpu-tiles 0 --e2e-chunks $chunks --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2>&1; echo "$name rc $?"; grep -E "hq_gen/deconv|hq_enc/conv2|rd/conv2" gpurun_out/layers_$name.csv | cut -d, -f1,2 | tr '\n' ' '; echo; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('   value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'layers', round(d['layers_ms_per_step'],3))
"; }
run def auto A=1
run c4 8,16,32,8 A=1
run c5 4,8,16,28,8 A=1
run c6 4,12,40,8 A=1
run c7 8,24,24,8 A=1
run noraster 8,16,32,8 CIC_TC_RASTER=0
