#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2o_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/r2o_layers.csv > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2o_bench.err
python - <<'PY'
import json,csv
d=json.load(open('gpurun_out/r2o_bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3), d['clocks'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'roof frac', round(e['frac_of_roof'],3))
print('roofline', d['roofline']['kernel'][:30], round(d['roofline']['frac'],3), {k:round(v['frac_of_hbm_peak'],3) for k,v in d['roofline']['hbm_kernels'].items()})
for n,c in d['configs'].items():
    if 'error' in c: print(n, c); continue
    print(n, 'value', round(c['value'],1), 'ms', round(c['ms_per_step'],3), 'e2e', round(c['e2e']['value'],1), 'parity', {k:v for k,v in (c.get('parity') or {}).items() if k in ('symbol_mismatches','recon_max_abs_01','psnr_delta_db_max','u8_max_abs_lsb','hq_ratio_delta_max')})
rows=list(csv.DictReader(open('gpurun_out/r2o_layers.csv')))
print(' '.join(f"{r['layer']}={float(r['ms']):.3f}" for r in rows if float(r['ms'])>0.2))
PY
