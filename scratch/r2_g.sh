#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2g_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_n2.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'], d['clocks'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'floor', round(e['host_copy_floor_ms'],2), 'roof frac', round(e['frac_of_roof'],3), 'host GB/s', round(e['host_gbs_achieved'],1))
for n,c in d['configs'].items():
    if 'error' in c: print(n, c); continue
    print(n, c['scaling'], 'value', round(c['value'],1), 'ms', round(c['ms_per_step'],3), 'e2e', round(c['e2e']['value'],1), c['e2e'].get('frac_of_roof'), c['config'].get('images_per_gpu'))
PY
