run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; grep -E "deconv4" gpurun_out/layers_$name.csv | cut -d, -f2 | tr '\n' ' '; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print(' ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])
"; }
for rep in 1 2; do
run a_m1nw4_$rep CIC_TC_MERGE=1 CIC_TC_NW=4
run b_m0nw4_$rep CIC_TC_MERGE=0 CIC_TC_NW=4
run c_m1nw2_$rep CIC_TC_MERGE=1 CIC_TC_NW=2
run d_m0nw2_$rep CIC_TC_MERGE=0 CIC_TC_NW=2
run e_m1nw1_$rep CIC_TC_MERGE=1 CIC_TC_NW=1
done
