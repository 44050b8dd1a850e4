#!/bin/bash
# usage: bash scratch/r2_scale.sh N   (default bench under torchrun on N GPUs)
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/r02c_bench_n$N.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02c_bench_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02c_bench_n{n}.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],3), d['clocks'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'floor', round(e['host_copy_floor_ms'],2), 'roof frac', round(e['frac_of_roof'],3), 'host GB/s', round(e['host_gbs_achieved'],1))
for c,v in d['configs'].items():
    if 'error' in v: print(c, v); continue
    print(c, v['scaling'], 'value', round(v['value'],1), 'ms', round(v['ms_per_step'],3), 'e2e', round(v['e2e']['value'],1), v['e2e'].get('frac_of_roof'), v['config'].get('images_per_gpu'))
PY
