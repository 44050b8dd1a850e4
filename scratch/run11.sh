run() { name=$1; chunks=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --cpu-tiles 0 --e2e-chunks $chunks --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo "$name rc $?"; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('   value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'layers', round(d['layers_ms_per_step'],3), d['clocks'], d['quality'])
"; tail -2 gpurun_out/err_$name.log; }
run g4 8,16,32,8 A=1
run g4t 8,16,32,8 CIC_PIPE_TIMELINE=1
run g3 8,48,8 A=1
run g5 4,8,16,28,8 A=1
run g6 4,4,8,16,24,8 A=1
run nog 8,16,32,8 CIC_PIPE_GRAPHS=0
