#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2e_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/r2e_layers.csv > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'], d['clocks'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'floor', round(e['host_copy_floor_ms'],2), 'roof frac', round(e['frac_of_roof'],3))
print(json.dumps(d['quality'])[:1500])
for n,c in d['configs'].items():
    if 'error' in c: print(n, c); continue
    print(n, 'value', round(c['value'],1), 'ms', round(c['ms_per_step'],3), 'e2e', round(c['e2e']['value'],1), c['e2e'].get('frac_of_roof'), json.dumps(c.get('quality'))[:300])
PY
