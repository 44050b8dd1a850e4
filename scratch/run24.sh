timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; grep -E "enc_conv1|hq_gen/deconv[34]|hq_enc/conv3" gpurun_out/layers_$name.csv | cut -d, -f2 | tr '\n' ' '; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print(' ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['quality']['psnr_db'])
"; tail -3 gpurun_out/err_$name.log; }
for rep in 1 2 3; do
run qt_$rep CIC_X=0
done
